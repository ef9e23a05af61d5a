// CChinHuhPermanentCalculator.h -- takes the place of piquassoboost/sampling/source/CChinHuhPermanentCalculator.h on the include
// path: the classes it declares are served by the B200 engine (see pqb200_pic_calculators.h).
#include "pqb200_pic_calculators.h"
