// selftest.cu -- TEST INFRASTRUCTURE ONLY: positive / negative controls of the emulation's racecheck (ThreadSanitizer build).
//   racy_kernel      reads a neighbour's shared word with no barrier after the write  -> must be reported
//   racy_warp_kernel the same inside one warp without __syncwarp                      -> must be reported
//   clean_kernel     the same exchange with __syncthreads / a shuffle                 -> must not be reported
// Also checks the execution model itself: barrier, shuffle, ballot, atomics.  Usage: selftest racy|racy_warp|clean
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>

__global__ void racy_kernel(int *out)
{
    __shared__ int s[64];
    s[threadIdx.x] = threadIdx.x * 3;
    out[threadIdx.x] = s[threadIdx.x ^ 32];          // the other warp's word: no __syncthreads
}

__global__ void racy_warp_kernel(int *out)
{
    __shared__ int s[32];
    s[threadIdx.x] = threadIdx.x * 3;
    out[threadIdx.x] = s[threadIdx.x ^ 1];           // a neighbouring lane's word: no __syncwarp
}

__global__ void clean_kernel(int *out, int *counter)
{
    __shared__ int s[64];
    s[threadIdx.x] = threadIdx.x * 3;
    __syncthreads();
    int v = s[threadIdx.x ^ 32];
    v += __shfl_xor_sync(0xffffffffu, (int)threadIdx.x, 1);
    v += __popc(__ballot_sync(0xffffffffu, threadIdx.x & 1));
    atomicAdd(counter, 1);
    out[threadIdx.x] = v;
}

int main(int argc, char **argv)
{
    int *out, *counter;
    cudaMalloc(&out, 64 * sizeof(int));
    cudaMalloc(&counter, sizeof(int));
    cudaMemset(counter, 0, sizeof(int));
    const char *what = argc > 1 ? argv[1] : "clean";
    if (!strcmp(what, "racy")) racy_kernel<<<2, 64>>>(out);
    else if (!strcmp(what, "racy_warp")) racy_warp_kernel<<<1, 32>>>(out);
    else {
        clean_kernel<<<3, 64>>>(out, counter);
        int h[64], c = 0;
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        cudaMemcpy(&c, counter, sizeof(c), cudaMemcpyDeviceToHost);
        for (int t = 0; t < 64; t++)
            if (h[t] != (t ^ 32) * 3 + (t ^ 1) + 16) { printf("selftest: wrong value at %d: %d\n", t, h[t]); return 1; }
        if (c != 192) { printf("selftest: counter %d\n", c); return 1; }
    }
    printf("selftest %s finished\n", what);
    return 0;
}
