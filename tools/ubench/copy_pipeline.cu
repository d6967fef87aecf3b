// copy_pipeline.cu -- what the host<->device link does under the access pattern of rtb_trace_host: chunk c goes H2D,
// (a small kernel), D2H on stream c % k, pinned buffers.  Prints single-copy times and the per-chunk timeline.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o copy_pipeline copy_pipeline.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__global__ void touch(double *p, size_t n) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; if (i < n) p[i] += 1.0; }
int main(int argc, char **argv)
{
    const size_t total = 1ull << 30;
    const size_t chunk = (argc > 1 ? atol(argv[1]) : 16) << 20;
    const int k = argc > 2 ? atoi(argv[2]) : 3;
    const int use_kernel = argc > 3 ? atoi(argv[3]) : 1;  // 0: no kernel between the copies
    const int gate = argc > 4 ? atoi(argv[4]) : 1;        // 0: enqueue everything up front (stream order alone protects the buffers)
    const int verbose = argc > 5 ? atoi(argv[5]) : 0;
    const int lag = argc > 6 ? atoi(argv[6]) : 0;         // 1: the copy-out of chunk c is issued behind the copy-in of chunk c + 1
    char *h_in, *h_out;
    CK(cudaHostAlloc(&h_in, total, cudaHostAllocDefault));
    CK(cudaHostAlloc(&h_out, total, cudaHostAllocDefault));
    for (size_t i = 0; i < total; i += 4096) h_in[i] = 1;
    std::vector<char *> d(k);
    std::vector<cudaStream_t> st(k);
    std::vector<cudaEvent_t> done(k);
    for (int s = 0; s < k; s++) { CK(cudaMalloc(&d[s], chunk)); CK(cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&done[s], cudaEventDisableTiming)); }
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float ms;
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaEventRecord(a, st[0])); CK(cudaMemcpyAsync(d[0], h_in, chunk, cudaMemcpyHostToDevice, st[0])); CK(cudaEventRecord(b, st[0]));
        CK(cudaStreamSynchronize(st[0])); CK(cudaEventElapsedTime(&ms, a, b));
        printf("H2D %zu MiB alone: %.3f ms  %.1f GB/s\n", chunk >> 20, ms, chunk / ms / 1e6);
        CK(cudaEventRecord(a, st[0])); CK(cudaMemcpyAsync(h_out, d[0], chunk, cudaMemcpyDeviceToHost, st[0])); CK(cudaEventRecord(b, st[0]));
        CK(cudaStreamSynchronize(st[0])); CK(cudaEventElapsedTime(&ms, a, b));
        printf("D2H %zu MiB alone: %.3f ms  %.1f GB/s\n", chunk >> 20, ms, chunk / ms / 1e6);
    }
    const size_t n_chunks = total / chunk;
    for (int rep = 0; rep < 3; rep++) {
        std::vector<cudaEvent_t> marks;
        auto mark = [&](cudaStream_t s) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); marks.push_back(e); };
        std::vector<bool> active(k, false);
        CK(cudaDeviceSynchronize());
        cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1);
        CK(cudaEventRecord(t0, st[0]));
        std::vector<cudaEvent_t> hdone(k);
        for (int s = 0; s < k; s++) CK(cudaEventCreateWithFlags(&hdone[s], cudaEventDisableTiming));
        for (size_t c = 0; c < n_chunks; c++) {
            const int s = (int)(c % k);
            if (gate && active[s]) CK(cudaEventSynchronize(done[s]));
            if (verbose && rep == 2) mark(st[s]);
            CK(cudaMemcpyAsync(d[s], h_in + c * chunk, chunk, cudaMemcpyHostToDevice, st[s]));
            if (verbose && rep == 2) mark(st[s]);
            if (lag) CK(cudaEventRecord(hdone[s], st[s]));
            if (use_kernel) touch<<<(unsigned)((chunk / 8 + 255) / 256), 256, 0, st[s]>>>((double *)d[s], chunk / 8);
            if (verbose && rep == 2) mark(st[s]);
            if (!lag) {
                CK(cudaMemcpyAsync(h_out + c * chunk, d[s], chunk, cudaMemcpyDeviceToHost, st[s]));
                if (verbose && rep == 2) mark(st[s]);
                CK(cudaEventRecord(done[s], st[s]));
                active[s] = true;
            } else if (c > 0) {
                const int p = (int)((c - 1) % k);
                CK(cudaStreamWaitEvent(st[p], hdone[s], 0));
                CK(cudaMemcpyAsync(h_out + (c - 1) * chunk, d[p], chunk, cudaMemcpyDeviceToHost, st[p]));
                CK(cudaEventRecord(done[p], st[p]));
                active[p] = true;
            }
        }
        if (lag) {
            const int p = (int)((n_chunks - 1) % k);
            CK(cudaMemcpyAsync(h_out + (n_chunks - 1) * chunk, d[p], chunk, cudaMemcpyDeviceToHost, st[p]));
        }
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(t1, st[0])); CK(cudaEventSynchronize(t1));
        CK(cudaEventElapsedTime(&ms, t0, t1));
        printf("pipeline chunk %zu MiB, %d streams, kernel %d, gate %d, lag %d: %.2f ms, %.1f GB/s\n", chunk >> 20, k, use_kernel, gate, lag, ms, 2.0 * total / ms / 1e6);
        for (size_t q = 0; q + 3 < marks.size() && q < 4 * 12; q += 4) {
            float x[4];
            for (int j = 0; j < 4; j++) cudaEventElapsedTime(&x[j], marks[0], marks[q + j]);
            printf("  chunk %2zu: in %7.3f - %7.3f, kernel done %7.3f, out done %7.3f\n", q / 4, x[0], x[1], x[2], x[3]);
        }
    }
    return 0;
}
