// fp64_switch.cu -- which FP64 instructions pay the B200's "warp switch" cycle?
// fp64_issue.cu found that a dependent DFMA chain per warp tops out at 66 % of the pipe however many warps are resident
// (3 cycles per instruction instead of 2).  This probe repeats the one-/two-/three-chain measurement for DADD, DMUL,
// DFMA with an immediate addend, DFMA with one distinct register, and a DMUL->DADD chain (the un-fused multiply-add the
// exact trace is made of), to tell an operand-delivery limit (third 64-bit register operand) from a per-warp issue limit.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_switch fp64_switch.cu
#include <cstdio>
#include <cuda_runtime.h>

// OP 0: x = fma(x, a, b)   1: x = x + a   2: x = x * a   3: x = fma(x, a, 1.0)   4: x = fma(x, x, x)
//    5: x = x * a; x = x + b (counted as two)   6: x = fma(a, b, x) (accumulate form)
template <int OP, int CHAINS>
__global__ void probe(double *out, int iters, double a, double b, long long *cycles)
{
    double x[CHAINS];
    for (int c = 0; c < CHAINS; c++) x[c] = threadIdx.x * 1e-9 + c;
    // per-thread values: the operands live in vector registers, as in real code (not in uniform registers)
    a += (threadIdx.x & 3) * 1e-15;
    b += (threadIdx.x & 5) * 1e-15;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int c = 0; c < CHAINS; c++) {
                if (OP == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(a), "d"(b));
                if (OP == 1) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(a));
                if (OP == 2) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(a));
                if (OP == 3) asm volatile("fma.rn.f64 %0, %0, %1, 0d3FF0000000000000;" : "+d"(x[c]) : "d"(a));
                if (OP == 4) asm volatile("fma.rn.f64 %0, %0, %0, %0;" : "+d"(x[c]));
                if (OP == 6) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(x[c]) : "d"(a), "d"(b));
            }
            if (OP == 5) {
#pragma unroll
                for (int c = 0; c < CHAINS; c++) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(a));
#pragma unroll
                for (int c = 0; c < CHAINS; c++) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(b));
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < CHAINS; c++) s += x[c];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int OP, int CHAINS>
void run(const char *what, int warps_per_smsp)
{
    double *out;
    long long *cyc;
    cudaMalloc(&out, 8);
    cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int r = 0; r < 2; r++) probe<OP, CHAINS><<<148, 128 * warps_per_smsp>>>(out, iters, 1.0000001, 1e-9, cyc);
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double fp64 = (double)iters * 16 * CHAINS * warps_per_smsp * (OP == 5 ? 2 : 1);
    printf("%-28s chains %d warps/smsp %d : %.3f instr/cyc/smsp (pipe %.0f%%), %.2f cycles per instruction\n", what, CHAINS,
           warps_per_smsp, fp64 / h, 200 * fp64 / h, h / fp64);
    cudaFree(out);
    cudaFree(cyc);
}

template <int OP>
void sweep(const char *what)
{
    for (int w : {1, 4, 7}) run<OP, 1>(what, w);
    for (int w : {4, 7}) run<OP, 2>(what, w);
    run<OP, 3>(what, 7);
    run<OP, 4>(what, 7);
}

int main()
{
    sweep<0>("DFMA x,a,b");
    sweep<1>("DADD x,a");
    sweep<2>("DMUL x,a");
    sweep<3>("DFMA x,a,1.0");
    sweep<4>("DFMA x,x,x");
    sweep<6>("DFMA a,b,x");
    sweep<5>("DMUL x,a ; DADD x,b");
    return 0;
}
