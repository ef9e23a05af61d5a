// CGeneralizedCliffordsBSimulationStrategy.h -- takes the place of
// piquassoboost/sampling/simulation_strategies/source/CGeneralizedCliffordsBSimulationStrategy.h on the include path:
// the class it declares is served by the B200 engine (see pqb200_pic_strategies.h).
#include "pqb200_pic_strategies.h"
