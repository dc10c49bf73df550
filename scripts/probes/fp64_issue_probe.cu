// Micro-benchmark: can other instructions issue in the gaps of a saturated FP64 pipe on B200?
// Each thread runs 8 independent DFMA chains plus K independent integer (IMAD) / FP32 (FFMA) /
// select (FSEL) operations per 8 DFMAs.  If the FP64 pipe only blocks FP64 issue slots, DFMA/s is
// unchanged as K grows (until the issue port saturates); if an FP64 instruction holds the
// scheduler's issue port for both of its cycles, time grows as 2*n_fp64 + n_other.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue_probe fp64_issue_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int KIND, int K>   // KIND 0: IMAD, 1: FFMA, 2: FSEL-like (fp32 select), 3: LDS
__global__ void probe(double* out, int* iout, int iters, int seed) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double x = 1.0000001, y = 1e-9;
    int i0 = threadIdx.x + seed, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 4, i5 = i0 + 5,
        i6 = i0 + 6, i7 = i0 + 7;
    float f0 = i0 * 1e-3f, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5,
          f6 = f0 + 6, f7 = f0 + 7;
    __shared__ int sh[256];
    sh[threadIdx.x & 255] = seed;
    __syncthreads();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
        a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (KIND == 0) {
                i0 = i0 * 3 + seed; i1 = i1 * 5 + seed; i2 = i2 * 7 + seed; i3 = i3 * 9 + seed;
                i4 = i4 * 11 + seed; i5 = i5 * 13 + seed; i6 = i6 * 15 + seed; i7 = i7 * 17 + seed;
            } else if (KIND == 1) {
                f0 = fmaf(f0, 1.0001f, 0.5f); f1 = fmaf(f1, 1.0001f, 0.5f);
                f2 = fmaf(f2, 1.0001f, 0.5f); f3 = fmaf(f3, 1.0001f, 0.5f);
                f4 = fmaf(f4, 1.0001f, 0.5f); f5 = fmaf(f5, 1.0001f, 0.5f);
                f6 = fmaf(f6, 1.0001f, 0.5f); f7 = fmaf(f7, 1.0001f, 0.5f);
            } else if (KIND == 2) {
                f0 = (i0 & 1) ? f1 : f0; f1 = (i0 & 2) ? f2 : f1; f2 = (i0 & 4) ? f3 : f2;
                f3 = (i0 & 8) ? f4 : f3; f4 = (i0 & 16) ? f5 : f4; f5 = (i0 & 32) ? f6 : f5;
                f6 = (i0 & 64) ? f7 : f6; f7 = (i0 & 128) ? f0 : f7;
                i0 = i0 + 1;
            } else {
                i0 += sh[(i0 + 0) & 255]; i1 += sh[(i1 + 1) & 255]; i2 += sh[(i2 + 2) & 255];
                i3 += sh[(i3 + 3) & 255]; i4 += sh[(i4 + 4) & 255]; i5 += sh[(i5 + 5) & 255];
                i6 += sh[(i6 + 6) & 255]; i7 += sh[(i7 + 7) & 255];
            }
        }
    }
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    iout[blockIdx.x * (size_t)blockDim.x + threadIdx.x] =
        i0 + i1 + i2 + i3 + i4 + i5 + i6 + i7 + (int)(f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7);
}

template <int KIND, int K>
void run(const char* name, int sms, int warps_per_sm) {
    const int threads = 256, blocks = sms * warps_per_sm / 8, iters = 1 << 14;
    double* d; int* di;
    cudaMalloc(&d, (size_t)threads * blocks * 8);
    cudaMalloc(&di, (size_t)threads * blocks * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        probe<KIND, K><<<blocks, threads>>>(d, di, iters, r);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    double dfma = (double)threads * blocks * 8.0 * iters / (best * 1e-3);
    double other = dfma * K;
    printf("{\"probe\":\"%s\",\"other_per_dfma\":%d,\"warps_per_sm\":%d,\"ms\":%.3f,\"dfma_per_s\":%.4g,"
           "\"other_per_s\":%.4g}\n", name, K, warps_per_sm, best, dfma, other);
    cudaFree(d); cudaFree(di);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    for (int w : {16, 32, 64}) {
        run<0, 0>("dfma_only", sms, w);
        run<0, 1>("imad", sms, w); run<0, 2>("imad", sms, w); run<0, 3>("imad", sms, w);
        run<1, 1>("ffma", sms, w); run<1, 2>("ffma", sms, w);
        run<2, 1>("fsel", sms, w); run<2, 2>("fsel", sms, w);
        run<3, 1>("lds", sms, w);
    }
    return 0;
}
