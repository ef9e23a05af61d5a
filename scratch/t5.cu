#include "../piquassoboost_b200/csrc/glynn_plain.cuh"
using namespace pqb200;
template __global__ void pqb200::glynn_plain_kernel<NN>(const __grid_constant__ PlainParams<NN>);
