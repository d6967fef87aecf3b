// psf_kernels.cu -- pupil grid -> PSF as a dense complex contraction (the "ray fan -> pupil phase -> PSF" tail of
// BASELINE.json).
//
// The reference only has this at script level (scripts/2022_02_06_perfect_imaging_system_psf.py:90-105): scattered
// pupil phases -> scipy griddata -> exp(i phi) -> mask -> fftshift(fft2(ifftshift(.))) -> |.|^2.  Here the pupil
// function is what the trace kernel accumulated per cell, P[v,u] = sum_rays exp(i (phi - phi_ref)) (optionally divided
// by the ray count of the cell), and the field at M x M image-plane sample points is the separable matrix product
//     E = A P B^T,   B[k,u] = exp(-2 pi i fx_k x_u),  A[l,v] = exp(-2 pi i fy_l y_v),   PSF = |E|^2
// with x_u, y_v the cell centres and fx_k, fy_l = (k - (M-1)/2) * df spatial frequencies (for a lens of focal length f:
// image coordinate = lambda f fx).  A "zoomed DFT": any sampling and window, no padding; for odd G, M = G and
// df = 1/(G*cell) it is exactly fftshift(fft2(ifftshift(P))) (odd M puts a sample on the zero frequency).
// fp64 throughout: tcgen05 has no f64 kind, so this runs on the FP64 pipe (64x64 register-tiled complex GEMM).
#include <cmath>
#include <math_constants.h>

#include "rtb_device.cuh"

namespace rtb {

namespace {

// P = (sum cos + i sum sin) / count, zero where the cell is empty.  Planar output (re plane, im plane).
__global__ void normalize_pupil_kernel(const double *__restrict__ grid, long long cells, double *__restrict__ re,
                                       double *__restrict__ im)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride) {
        const double c = grid[2 * cells + i];
        const double inv = c > 0.0 ? 1.0 / c : 0.0;
        re[i] = grid[i] * inv;
        im[i] = grid[cells + i] * inv;
    }
}

// D[k, u] = exp(-2 pi i f_k x_u), f_k = (k - (M-1)/2) df, x_u = (u + 0.5) cell - half.  Planar (M x G).
__global__ void dft_matrix_kernel(int M, int G, double df, double cell, double half, double *__restrict__ re,
                                  double *__restrict__ im)
{
    const long long n = (long long)M * G;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int k = (int)(i / G), u = (int)(i % G);
        const double f = ((double)k - 0.5 * (double)(M - 1)) * df;
        const double x = ((double)u + 0.5) * cell - half;
        double s, c;
        sincospi(-2.0 * f * x, &s, &c);
        re[i] = c;
        im[i] = s;
    }
}

// C[m, n] = sum_k X[m, k] * Y[n, k]   (complex, planar, row-major; X is Mx x K, Y is Ny x K)
// abs2 != 0: write |C|^2 into c_re only.  64 x 64 tile per block, 16 x 16 threads, 4 x 4 complex accumulators each.
constexpr int kTile = 64, kStep = 16;
__global__ void __launch_bounds__(256) zgemm_nt_kernel(int Mx, int Ny, int K, const double *__restrict__ x_re,
                                                       const double *__restrict__ x_im, const double *__restrict__ y_re,
                                                       const double *__restrict__ y_im, double *__restrict__ c_re,
                                                       double *__restrict__ c_im, int abs2)
{
    __shared__ double sxr[kStep][kTile + 1], sxi[kStep][kTile + 1], syr[kStep][kTile + 1], syi[kStep][kTile + 1];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int m0 = blockIdx.y * kTile, n0 = blockIdx.x * kTile;
    double ar[4][4] = {}, ai[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += kStep) {
        // stage 64 x 16 of X and of Y (k fastest in memory -> coalesced along k)
        for (int e = threadIdx.x; e < kTile * kStep; e += 256) {
            const int r = e / kStep, kk = e % kStep;
            const int k = k0 + kk;
            const bool kin = k < K;
            const int m = m0 + r, n = n0 + r;
            sxr[kk][r] = (kin && m < Mx) ? x_re[(long long)m * K + k] : 0.0;
            sxi[kk][r] = (kin && m < Mx) ? x_im[(long long)m * K + k] : 0.0;
            syr[kk][r] = (kin && n < Ny) ? y_re[(long long)n * K + k] : 0.0;
            syi[kk][r] = (kin && n < Ny) ? y_im[(long long)n * K + k] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kStep; kk++) {
            double xr[4], xi[4], yr[4], yi[4];
#pragma unroll
            for (int a = 0; a < 4; a++) {
                xr[a] = sxr[kk][ty * 4 + a];
                xi[a] = sxi[kk][ty * 4 + a];
                yr[a] = syr[kk][tx * 4 + a];
                yi[a] = syi[kk][tx * 4 + a];
            }
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    ar[a][b] = fma(xr[a], yr[b], fma(-xi[a], yi[b], ar[a][b]));
                    ai[a][b] = fma(xr[a], yi[b], fma(xi[a], yr[b], ai[a][b]));
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int m = m0 + ty * 4 + a, n = n0 + tx * 4 + b;
            if (m < Mx && n < Ny) {
                const long long o = (long long)m * Ny + n;
                if (abs2) {
                    c_re[o] = fma(ar[a][b], ar[a][b], ai[a][b] * ai[a][b]);
                } else {
                    c_re[o] = ar[a][b];
                    c_im[o] = ai[a][b];
                }
            }
        }
}

} // namespace

// scratch layout (doubles): [0, 2MG) DFT matrix D (re, im) ; [2MG, 4MG) Wt (re, im) ; [4MG, 4MG + 2GG) normalised pupil
long long psf_scratch_doubles(int G, int M, int normalize)
{
    return 4LL * M * G + (normalize ? 2LL * G * G : 0LL);
}

cudaError_t launch_psf(const double *grid, int G, double half_width, int M, double df, int normalize, double *scratch,
                       double *psf_out, double *field_re, double *field_im, int sm_count, cudaStream_t stream, int *launches)
{
    const long long cells = (long long)G * G, mg = (long long)M * G;
    double *d_re = scratch, *d_im = scratch + mg, *w_re = scratch + 2 * mg, *w_im = scratch + 3 * mg;
    const double *p_re = grid, *p_im = grid + cells;
    int n = 0;
    if (normalize) {
        double *q_re = scratch + 4 * mg, *q_im = q_re + cells;
        normalize_pupil_kernel<<<sm_count * 8, 256, 0, stream>>>(grid, cells, q_re, q_im);
        p_re = q_re;
        p_im = q_im;
        n++;
    }
    const double cell = 2.0 * half_width / (double)G;
    dft_matrix_kernel<<<sm_count * 8, 256, 0, stream>>>(M, G, df, cell, half_width, d_re, d_im);
    n++;
    // Wt[k, v] = sum_u D[k, u] P[v, u]
    dim3 g1((G + kTile - 1) / kTile, (M + kTile - 1) / kTile);
    zgemm_nt_kernel<<<g1, 256, 0, stream>>>(M, G, G, d_re, d_im, p_re, p_im, w_re, w_im, 0);
    n++;
    // E[l, k] = sum_v D[l, v] Wt[k, v]   (square pupil grid: the same DFT matrix serves both axes)
    dim3 g2((M + kTile - 1) / kTile, (M + kTile - 1) / kTile);
    if (field_re && field_im) {
        zgemm_nt_kernel<<<g2, 256, 0, stream>>>(M, M, G, d_re, d_im, w_re, w_im, field_re, field_im, 0);
        n++;
    }
    if (psf_out) {
        zgemm_nt_kernel<<<g2, 256, 0, stream>>>(M, M, G, d_re, d_im, w_re, w_im, psf_out, nullptr, 1);
        n++;
    }
    if (launches) *launches = n;
    return cudaGetLastError();
}

} // namespace rtb
