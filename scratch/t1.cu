#include "../piquassoboost_b200/csrc/glynn_plain.cuh"
using namespace pqb200;
#ifndef NN
#define NN 40
#endif
#ifndef SS
#define SS 2
#endif
template __global__ void pqb200::glynn_plain_kernel<NN,SS>(const __grid_constant__ PlainParams<NN,SS>);
