#include "v3.cuh"
using namespace pqb200;
#ifndef NN
#define NN 40
#endif
template __global__ void pqb200::glynn_u_kernel<NN,NN>(const __grid_constant__ UParams<NN,NN>);
