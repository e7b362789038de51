// smem_stream.cu — what fraction of the nominal 128 B/clk/SM does an in-place LDS.64 + STS.64 stream reach?
// (denominator sanity check for roofline_smem in bench.py).  nvcc -arch=sm_100a -O3 -o smem_stream smem_stream.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>   // 0: load+store in place (what BP does), 1: loads only, 2: stores only
__global__ void __launch_bounds__(1024, 1) stream(double *out, int iters, int words)
{
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < words; i += blockDim.x) sm[i] = (double)i;
    __syncthreads();
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
        for (int i = threadIdx.x; i < words; i += blockDim.x) {
            if (MODE == 0) { sm[i] = sm[i] + 1.0; }
            else if (MODE == 1) { acc += sm[i]; }
            else { sm[i] = acc + (double)it; }
        }
        __syncthreads();
    }
    if (MODE == 1) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    else out[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x];
}

template <int MODE>
static void run(const char *name, int threads)
{
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const int words = 24 * 1024;                 // 192 KB of doubles
    const size_t smem = (size_t)words * 8;
    cudaFuncSetAttribute(stream<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    double *out;
    cudaMalloc(&out, sizeof(double) * sms * 1024);
    const int iters = 2000;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    stream<MODE><<<sms, threads, smem>>>(out, 10, words);
    cudaEventRecord(a);
    stream<MODE><<<sms, threads, smem>>>(out, iters, words);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double bytes = (double)iters * words * 8 * (MODE == 0 ? 2 : 1) * sms;
    const double gbs = bytes / (ms * 1e-3) / 1e9;
    const double peak = (double)sms * 128 * khz * 1e3 / 1e9;
    printf("%-22s threads=%4d  %8.1f GB/s  = %.3f of %d SMs x 128 B/clk x %.0f MHz (%.1f GB/s)\n", name, threads, gbs,
           gbs / peak, sms, khz / 1e3, peak);
    cudaFree(out);
}

int main()
{
    for (int t : {512, 800, 1024}) {
        run<0>("ld.64 + st.64 in place", t);
        run<1>("ld.64 only", t);
        run<2>("st.64 only", t);
    }
    return 0;
}
