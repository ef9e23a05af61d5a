#include "v2.cuh"
using namespace pqb200;
#ifndef NN
#define NN 40
#endif
#ifndef MODE_
#define MODE_ 0
#endif
template __global__ void pqb200::glynn_rolled_kernel<NN,MODE_,NN>(const __grid_constant__ RolledParams<NN,NN>);
