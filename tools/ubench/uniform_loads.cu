// uniform_loads.cu -- when does ptxas (12.9, sm_100a) read a dynamically indexed kernel-parameter array through the
// uniform datapath (LDCU.64 UR, c[0x0][UR+off]; FP64 instructions then take the value as a uniform-register operand,
// which costs no operand-port cycle) instead of LDC into vector registers?  Compile-only experiment behind DESIGN.md 4a:
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -cubin -o uniform_loads.cubin uniform_loads.cu
//     cuobjdump -sass uniform_loads.cubin | grep -E "Function|LDCU.*\[UR|LDC.*\[R"
// Result (count of indexed LDCU / indexed vector LDC per kernel):
//   kA  loop, alive flag combined with `&`                         60 / 0    uniform
//   kA' the same with `&&` (a predicated constant load appears)     0 / 60   vector
//   kB  + vote-and-break                                            4 / 0    uniform
//   kC  + outer grid-stride loop            kD + guarded load/store 60 / 0    uniform
//   kE  branch on P.s[k].kind inside the loop                       0 / 35   vector
//   kF  branch on a shared-memory code inside the loop              0 / 28   vector
//   kG / kH  the same branches taken through __all_sync             0 / ..   vector
//   kL  runs from kernel parameters, one tight loop per run        77 / 0    uniform
//   kK  runs whose code comes from shared memory                    0 / 75   vector
//   kO / kP  a top-level branch on a flag read from memory          7 / 28, 0 / 70   vector
//   kQ / kR / kT  kL + shared arrays indexed by k, per-lane shared pointers, an out-of-line call after the loop   uniform
// i.e. the loop index stays uniform only while NO branch of the kernel depends on a value loaded from memory (or on a
// per-lane predicated constant load); one such branch anywhere -- even in another loop with its own index -- makes ptxas
// keep every index in a vector register.  The lean kernel's probe-driven dispatch is such a branch.
struct S { double a, b, c, d; int kind; int pad; };
struct Params { int n; int nruns; S s[64]; int run_end[64]; int run_code[64]; };
// A: t1 + alive flag (no vote)
__global__ void kA(const __grid_constant__ Params P, double *out, const double *in) {
  double x = in[threadIdx.x];
  bool alive = x > 0;
  for (int k = 0; k < P.n; k++) { x = x * P.s[k].a + P.s[k].b; x = x * x - P.s[k].c; alive = alive & (x < P.s[k].d); }
  out[threadIdx.x] = alive ? x : 0.0;
}
// B: A + vote break
__global__ void kB(const __grid_constant__ Params P, double *out, const double *in) {
  double x = in[threadIdx.x];
  bool alive = x > 0;
  for (int k = 0; k < P.n; k++) { if (!__any_sync(0xffffffffu, alive)) break; x = x * P.s[k].a + P.s[k].b; x = x * x - P.s[k].c; alive = alive & (x < P.s[k].d); }
  out[threadIdx.x] = alive ? x : 0.0;
}
// C: A inside uniform outer loop
__global__ void kC(const __grid_constant__ Params P, double *out, const double *in, int n) {
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
  int i = base + threadIdx.x;
  double x = in[i];
  bool alive = x > 0;
  for (int k = 0; k < P.n; k++) { x = x * P.s[k].a + P.s[k].b; x = x * x - P.s[k].c; alive = alive & (x < P.s[k].d); }
  out[i] = alive ? x : 0.0;
  }
}
// D: C with guarded load/store
__global__ void kD(const __grid_constant__ Params P, double *out, const double *in, int n) {
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
  int i = base + threadIdx.x;
  double x = 0; if (i < n) x = in[i];
  bool alive = x > 0;
  for (int k = 0; k < P.n; k++) { x = x * P.s[k].a + P.s[k].b; x = x * x - P.s[k].c; alive = alive & (x < P.s[k].d); }
  if (i < n) out[i] = alive ? x : 0.0;
  }
}
// E: A with kind switch (uniform from const)
__global__ void kE(const __grid_constant__ Params P, double *out, const double *in) {
  double x = in[threadIdx.x];
  bool alive = x > 0;
  for (int k = 0; k < P.n; k++) { if (P.s[k].kind == 1) { x = x * P.s[k].a + P.s[k].b; } else { x = x * x - P.s[k].c; } alive = alive & (x < P.s[k].d); }
  out[threadIdx.x] = alive ? x : 0.0;
}
// F: A with a shared-memory code switch
__global__ void kF(const __grid_constant__ Params P, double *out, const double *in) {
  __shared__ int code[64];
  if (threadIdx.x < 64) code[threadIdx.x] = in[threadIdx.x] > 1 ? 1 : 0;
  __syncthreads();
  double x = in[threadIdx.x];
  bool alive = x > 0;
  for (int k = 0; k < P.n; k++) { if (code[k] == 1) { x = x * P.s[k].a + P.s[k].b; } else { x = x * x - P.s[k].c; } alive = alive & (x < P.s[k].d); }
  out[threadIdx.x] = alive ? x : 0.0;
}
// G: shared code switch through votes
__global__ void kG(const __grid_constant__ Params P, double *out, const double *in) {
  __shared__ int code[64];
  if (threadIdx.x < 64) code[threadIdx.x] = in[threadIdx.x] > 1 ? 1 : 0;
  __syncthreads();
  double x = in[threadIdx.x];
  bool alive = x > 0;
  for (int k = 0; k < P.n; k++) { if (__all_sync(0xffffffffu, code[k] == 1)) { x = x * P.s[k].a + P.s[k].b; } else { x = x * x - P.s[k].c; } alive = alive & (x < P.s[k].d); }
  out[threadIdx.x] = alive ? x : 0.0;
}
// H: constant kind switch through votes
__global__ void kH(const __grid_constant__ Params P, double *out, const double *in) {
  double x = in[threadIdx.x];
  bool alive = x > 0;
  for (int k = 0; k < P.n; k++) { if (__all_sync(0xffffffffu, P.s[k].kind == 1)) { x = x * P.s[k].a + P.s[k].b; } else { x = x * x - P.s[k].c; } alive = alive & (x < P.s[k].d); }
  out[threadIdx.x] = alive ? x : 0.0;
}
// I: G plus a per-lane divergent block after the uniform part (like the reduction sample)
__global__ void kI(const __grid_constant__ Params P, double *out, const double *in, int kred) {
  __shared__ int code[64];
  if (threadIdx.x < 64) code[threadIdx.x] = in[threadIdx.x] > 1 ? 1 : 0;
  __syncthreads();
  double x = in[threadIdx.x];
  bool alive = x > 0;
  for (int k = 0; k < P.n; k++) { if (__all_sync(0xffffffffu, code[k] == 1)) { x = x * P.s[k].a + P.s[k].b; } else { x = x * x - P.s[k].c; } alive = alive & (x < P.s[k].d);
     if (k == kred && alive) atomicAdd(out + 1000, x); }
  out[threadIdx.x] = alive ? x : 0.0;
}
// J: like I but the divergent block guarded by a vote first
__global__ void kJ(const __grid_constant__ Params P, double *out, const double *in, int kred) {
  __shared__ int code[64];
  if (threadIdx.x < 64) code[threadIdx.x] = in[threadIdx.x] > 1 ? 1 : 0;
  __syncthreads();
  double x = in[threadIdx.x];
  bool alive = x > 0;
  for (int k = 0; k < P.n; k++) { if (__all_sync(0xffffffffu, code[k] == 1)) { x = x * P.s[k].a + P.s[k].b; } else { x = x * x - P.s[k].c; } alive = alive & (x < P.s[k].d);
     if (k == kred) { if (alive) atomicAdd(out + 1000, x); } }
  out[threadIdx.x] = alive ? x : 0.0;
}
// K: runs, code from shared via vote
__global__ void kK(const __grid_constant__ Params P, double *out, const double *in) {
  __shared__ int code[64]; __shared__ int rend[64];
  if (threadIdx.x < 64) { code[threadIdx.x] = in[threadIdx.x] > 1 ? 1 : 0; rend[threadIdx.x] = (int)in[64 + threadIdx.x]; }
  __syncthreads();
  double x = in[threadIdx.x];
  bool alive = x > 0;
  int k = 0;
  for (int r = 0; r < P.nruns; r++) {
    const int end = __shfl_sync(0xffffffffu, rend[r], 0);
    if (__all_sync(0xffffffffu, code[r] == 1)) { for (; k < end; k++) { x = x * P.s[k].a + P.s[k].b; alive = alive & (x < P.s[k].d); } }
    else { for (; k < end; k++) { x = x * x - P.s[k].c; alive = alive & (x < P.s[k].d);} }
  }
  out[threadIdx.x] = alive ? x : 0.0;
}
// L: runs from kernel params
__global__ void kL(const __grid_constant__ Params P, double *out, const double *in) {
  double x = in[threadIdx.x];
  bool alive = x > 0;
  int k = 0;
  for (int r = 0; r < P.nruns; r++) {
    const int end = P.run_end[r];
    if (P.run_code[r] == 1) { for (; k < end; k++) { x = x * P.s[k].a + P.s[k].b; alive = alive & (x < P.s[k].d); } }
    else { for (; k < end; k++) { x = x * x - P.s[k].c; alive = alive & (x < P.s[k].d);} }
  }
  out[threadIdx.x] = alive ? x : 0.0;
}
// M: single loop, if/else on a kernel-param-derived uniform condition not indexed by k
__global__ void kM(const __grid_constant__ Params P, double *out, const double *in) {
  double x = in[threadIdx.x];
  bool alive = x > 0;
  for (int k = 0; k < P.n; k++) { if (P.nruns == 1) { x = x * P.s[k].a + P.s[k].b; } else { x = x * x - P.s[k].c; } alive = alive & (x < P.s[k].d); }
  out[threadIdx.x] = alive ? x : 0.0;
}
// N: two sequential loops (no outer run loop): spheres then flats
__global__ void kN(const __grid_constant__ Params P, double *out, const double *in) {
  double x = in[threadIdx.x];
  bool alive = x > 0;
  int k = 0;
  for (; k < P.nruns; k++) { x = x * P.s[k].a + P.s[k].b; alive = alive & (x < P.s[k].d); }
  for (; k < P.n; k++) { x = x * x - P.s[k].c; alive = alive & (x < P.s[k].d); }
  out[threadIdx.x] = alive ? x : 0.0;
}
// O: top-level branch on a flag read from global memory; inside, run-structured loops from params
__global__ void kO(const __grid_constant__ Params P, double *out, const double *in, const int *flag) {
  __shared__ int code[64];
  __shared__ int sflag;
  if (threadIdx.x < 64) code[threadIdx.x] = in[threadIdx.x] > 1 ? 1 : 0;
  if (threadIdx.x == 0) sflag = flag[0];
  __syncthreads();
  for (int base = blockIdx.x * blockDim.x; base < 100000; base += gridDim.x * blockDim.x) {
  double x = in[base + threadIdx.x];
  bool alive = x > 0;
  if (sflag == 0) {
    int k = 0;
    for (int r = 0; r < P.nruns; r++) {
      const int end = P.run_end[r];
      if (P.run_code[r] == 1) { for (; k < end; k++) { if (!__any_sync(0xffffffffu, alive)) break; x = x * P.s[k].a + P.s[k].b; alive = alive & (x < P.s[k].d); } }
      else { for (; k < end; k++) { if (!__any_sync(0xffffffffu, alive)) break; x = x * x - P.s[k].c; alive = alive & (x < P.s[k].d);} }
    }
  } else {
    for (int k = 0; k < P.n; k++) { if (code[k] == 1) { x = x * P.s[k].a + P.s[k].b; } else { x = x * x - P.s[k].c; } alive = alive & (x < P.s[k].d); }
  }
  out[base + threadIdx.x] = alive ? x : 0.0;
  }
}
// P: the same with the flag through __syncthreads_or (block-uniform by construction)
__global__ void kP(const __grid_constant__ Params P, double *out, const double *in, const int *flag) {
  __shared__ int code[64];
  if (threadIdx.x < 64) code[threadIdx.x] = in[threadIdx.x] > 1 ? 1 : 0;
  const int f = __syncthreads_or(threadIdx.x == 0 ? flag[0] : 0);
  for (int base = blockIdx.x * blockDim.x; base < 100000; base += gridDim.x * blockDim.x) {
  double x = in[base + threadIdx.x];
  bool alive = x > 0;
  if (f == 0) {
    int k = 0;
    for (int r = 0; r < P.nruns; r++) {
      const int end = P.run_end[r];
      if (P.run_code[r] == 1) { for (; k < end; k++) { if (!__any_sync(0xffffffffu, alive)) break; x = x * P.s[k].a + P.s[k].b; alive = alive & (x < P.s[k].d); } }
      else { for (; k < end; k++) { if (!__any_sync(0xffffffffu, alive)) break; x = x * x - P.s[k].c; alive = alive & (x < P.s[k].d);} }
    }
  } else {
    for (int k = 0; k < P.n; k++) { if (code[k] == 1) { x = x * P.s[k].a + P.s[k].b; } else { x = x * x - P.s[k].c; } alive = alive & (x < P.s[k].d); }
  }
  out[base + threadIdx.x] = alive ? x : 0.0;
  }
}
// Q: kL + a shared array indexed by k
__global__ void kQ(const __grid_constant__ Params P, double *out, const double *in) {
  __shared__ double sh[64];
  if (threadIdx.x < 64) sh[threadIdx.x] = in[threadIdx.x];
  __syncthreads();
  double x = in[threadIdx.x];
  bool alive = x > 0;
  int k = 0;
  for (int r = 0; r < P.nruns; r++) {
    const int end = P.run_end[r];
    if (P.run_code[r] == 1) { for (; k < end; k++) { x = x * P.s[k].a + P.s[k].b + sh[k]; alive = alive & (x < P.s[k].d); } }
    else { for (; k < end; k++) { x = x * x - P.s[k].c; alive = alive & (x < P.s[k].d);} }
  }
  out[threadIdx.x] = alive ? x : 0.0;
}
// R: kL + a per-lane shared pointer that advances (like the index pairs)
__global__ void kR(const __grid_constant__ Params P, double *out, const double *in) {
  __shared__ double sh[512];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) sh[i] = in[i];
  __syncthreads();
  double x = in[threadIdx.x];
  const double *pair = sh + ((int)x & 3) * 64;
  bool alive = x > 0;
  int k = 0;
  for (int r = 0; r < P.nruns; r++) {
    const int end = P.run_end[r];
    if (P.run_code[r] == 1) { for (; k < end; k++) { x = x * P.s[k].a + P.s[k].b + pair[0]; pair += 2; alive = alive & (x < P.s[k].d); } }
    else { for (; k < end; k++) { x = x * x - P.s[k].c + pair[1]; pair += 2; alive = alive & (x < P.s[k].d);} }
  }
  out[threadIdx.x] = alive ? x : 0.0;
}
// T: kR + vote break + failed flag + out-of-line call after
__device__ __noinline__ double slow(const Params *P, double x) { for (int k = 0; k < P->n; k++) x = x / P->s[k].a; return x; }
__global__ void kT(const __grid_constant__ Params P, double *out, const double *in, int n) {
  __shared__ double sh[512];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) sh[i] = in[i];
  __syncthreads();
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
  const int i = base + threadIdx.x;
  const bool valid = i < n;
  double x = 0; if (valid) x = in[i];
  const double *pair = sh + ((int)x & 3) * 64;
  bool alive = valid & (x > 0);
  bool failed = false;
  int k = 0;
  for (int r = 0; r < P.nruns; r++) {
    const int end = P.run_end[r];
    if (P.run_code[r] == 1) { for (; k < end; k++) { if (!__any_sync(0xffffffffu, alive)) break; x = x * P.s[k].a + P.s[k].b + pair[0]; pair += 2; bool ok = x > P.s[k].c; failed = failed | (alive & !ok); alive = alive & ok & (x < P.s[k].d); } }
    else { for (; k < end; k++) { if (!__any_sync(0xffffffffu, alive)) break; x = x * x - P.s[k].c + pair[1]; pair += 2; alive = alive & (x < P.s[k].d);} }
  }
  if (failed) x = slow(&P, x);
  if (valid) out[i] = alive ? x : 0.0;
  }
}
