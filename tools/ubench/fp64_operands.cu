// fp64_operands.cu -- does the FP64 pipe's rate depend on how many distinct register operands an instruction reads?
// 8 independent chains per thread (plenty of ILP), 4 warps per scheduler; variants differ only in operand sources.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_operands fp64_operands.cu
#include <cstdio>
#include <cuda_runtime.h>

// KIND 0: x = fma(x, m, c)  m, c shared registers (what the DFMA peak probe does)
// KIND 1: x = fma(x, y_k, z_k)  three distinct register pairs per instruction
// KIND 2: x = x * y_k  (DMUL, two distinct)      KIND 3: x = x + y_k (DADD, two distinct)
// KIND 4: x_k = fma(x_k, x_{k+1}, x_{k+2}) all operands rotating through the chain registers (like real code)
template <int KIND>
__global__ void probe(double *out, int iters, double a, double b, long long *cycles)
{
    double x[8], y[8], z[8];
    for (int c = 0; c < 8; c++) {
        x[c] = threadIdx.x * 1e-9 + c;
        y[c] = a + c * 1e-12;
        z[c] = b + c * 1e-13;
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int c = 0; c < 8; c++) {
                if (KIND == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(a), "d"(b));
                if (KIND == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(y[c]), "d"(z[c]));
                if (KIND == 2) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(y[c]));
                if (KIND == 3) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(y[c]));
                if (KIND == 4) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(y[(c + u) & 7]), "d"(z[(c + 3 * u + 1) & 7]));
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < 8; c++) s += x[c] + y[c] + z[c];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int KIND>
void run(const char *what, int warps_per_smsp)
{
    double *out;
    long long *cyc;
    cudaMalloc(&out, 8);
    cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int r = 0; r < 2; r++) probe<KIND><<<148, 128 * warps_per_smsp>>>(out, iters, 1.0000001, 1e-9, cyc);
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double fp64 = (double)iters * 64 * warps_per_smsp;
    printf("%-44s warps/smsp %d : %.3f instr/cyc/smsp (pipe %.0f%%)\n", what, warps_per_smsp, fp64 / h, 200 * fp64 / h);
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    for (int w : {4, 7}) {
        run<0>("DFMA x,m,c (2 shared operands)", w);
        run<1>("DFMA x,y_k,z_k (3 distinct registers)", w);
        run<4>("DFMA x,y_j,z_l (3 distinct, rotating)", w);
        run<2>("DMUL x,y_k (2 distinct)", w);
        run<3>("DADD x,y_k (2 distinct)", w);
    }
    return 0;
}
