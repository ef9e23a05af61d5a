struct P { double v[64]; double* out; };
__global__ void k(const __grid_constant__ P p) {
    double x = p.out[threadIdx.x];
    double y = x*3.0;
    #pragma unroll
    for (int i=0;i<8;i++){ x += p.v[i]; y = fma(y, p.v[i+8], x); x -= p.v[i+16]; }
    p.out[threadIdx.x] = x+y;
}
