// strided_copy_probe.cu — what HBM gives a tile-wise copy whose rows are W bytes wide and `pitch` bytes apart (measurement only).
// Models the global-memory side of the column kernels without any FFT work: tile t = rows j = 0..NR-1 of W contiguous bytes at
// in + j * pitch + t * W, copied to the same place in out (or to a dense block: mode 1, or from a dense block: mode 2).
// One persistent CTA of 1024 threads per SM slot; every thread keeps 8 x 16-byte loads in flight.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o strided_copy_probe strided_copy_probe.cu && ./strided_copy_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024, 1) probe(const char* __restrict__ in, char* __restrict__ out, long long pitch, int W, int NR,
                                                 long long ntiles, int mode) {
    const int ppr = W / 16;                  // 16-byte pieces per row
    const int total = NR * ppr;              // pieces per tile
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long long tiles_per_row = pitch / W;
        const long long blk = t / tiles_per_row, col = t % tiles_per_row;     // tiles beyond one row block start a new block of NR rows
        const char* src0 = in + blk * NR * pitch + col * W;
        char* dst0 = out + blk * NR * pitch + col * W;
        const long long dense = t * (long long)NR * W;
        for (int p0 = threadIdx.x; p0 < total; p0 += 8 * 1024) {
            int4 v[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int p = p0 + m * 1024;
                if (p < total) {
                    const int j = p / ppr, part = p % ppr;
                    const char* s = mode == 2 ? in + dense + (long long)p * 16 : src0 + (long long)j * pitch + part * 16;
                    v[m] = *reinterpret_cast<const int4*>(s);
                }
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int p = p0 + m * 1024;
                if (p < total) {
                    const int j = p / ppr, part = p % ppr;
                    char* d = mode == 1 ? out + dense + (long long)p * 16 : dst0 + (long long)j * pitch + part * 16;
                    *reinterpret_cast<int4*>(d) = v[m];
                }
            }
        }
    }
}

int main() {
    const size_t bytes = 4ull << 30;          // 4 GiB in, 4 GiB out: far beyond L2
    char *in, *out;
    cudaMalloc(&in, bytes); cudaMalloc(&out, bytes);
    cudaMemset(in, 1, bytes); cudaMemset(out, 0, bytes);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Case { const char* name; long long pitch; int W, NR, mode, ctas; };
    const Case cases[] = {
        {"contiguous rows (pitch = W = 32 KiB), 1 CTA/SM", 32768, 32768, 4, 0, 1},
        {"32-byte pieces, 32 KiB pitch, 4096 rows per tile (c5b pass 1: both sides)", 32768, 32, 4096, 0, 1},
        {"32-byte pieces in, dense out", 32768, 32, 4096, 1, 1},
        {"dense in, 32-byte pieces out (c5b pass 2)", 32768, 32, 4096, 2, 1},
        {"64-byte pieces, 32 KiB pitch, 2048 rows", 32768, 64, 2048, 0, 1},
        {"128-byte pieces, 32 KiB pitch, 1024 rows", 32768, 128, 1024, 0, 1},
        {"256-byte pieces, 64 KiB pitch, 512 rows (c2 column passes)", 65536, 256, 512, 0, 1},
        {"16-byte pieces, 64 KiB pitch, 8192 rows (one-pass c2 columns)", 65536, 16, 8192, 0, 1},
        {"32-byte pieces, 32 KiB pitch, 4096 rows, 2 CTAs/SM", 32768, 32, 4096, 0, 2},
    };
    for (const Case& c : cases) {
        const long long tile_bytes = (long long)c.NR * c.W;
        const long long ntiles = (long long)(bytes / tile_bytes);
        const int grid = sms * c.ctas;
        for (int w = 0; w < 2; ++w) probe<<<grid, 1024>>>(in, out, c.pitch, c.W, c.NR, ntiles, c.mode);
        cudaEventRecord(e0);
        for (int r = 0; r < 3; ++r) probe<<<grid, 1024>>>(in, out, c.pitch, c.W, c.NR, ntiles, c.mode);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= 3;
        cudaError_t err = cudaGetLastError();
        printf("{\"case\": \"%s\", \"ms\": %.4f, \"GB/s\": %.1f, \"frac_of_copy_peak\": %.3f%s}\n", c.name, ms, 2.0 * ntiles * tile_bytes / (ms * 1e-3) / 1e9,
               2.0 * ntiles * tile_bytes / (ms * 1e-3) / 1e9 / 6547.8, err == cudaSuccess ? "" : ", \"error\": true");
    }
    return 0;
}
