// Microbenchmark: cost of remote NVLink stores + system fences + flag publication from one kernel (single process,
// cudaDeviceEnablePeerAccess).  nvcc -arch=sm_100a -O3 p2p_lat.cu -o p2p_lat && ./p2p_lat
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void push(const uint4* src, uint4* dst_local, uint4* dst_remote, int n16, int mode, unsigned long long* flag_remote,
                     unsigned int* done, unsigned long long seq) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) {
        uint4 v = src[i];
        if (mode >= 0) dst_local[i] = v;
        if (mode >= 1) dst_remote[i] = v;
    }
    if (mode >= 2) {
        __syncthreads();
        __shared__ bool last;
        if (threadIdx.x == 0) {
            if (mode == 2 || mode >= 4) __threadfence_system();
            if (mode == 3) __threadfence();
            last = atomicAdd(done, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (last && threadIdx.x == 0) {
            if (mode >= 4) {
                __threadfence_system();
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag_remote), "l"(seq) : "memory");
            }
            *done = 0;
        }
    }
}

int main() {
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
    CK(cudaSetDevice(1));
    uint4* remote; unsigned long long* flag_remote;
    CK(cudaMalloc(&remote, 1 << 20));
    CK(cudaMalloc(&flag_remote, 64));
    CK(cudaSetDevice(0));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    uint4 *src, *local; unsigned int* done;
    CK(cudaMalloc(&src, 1 << 20)); CK(cudaMalloc(&local, 1 << 20)); CK(cudaMalloc(&done, 64));
    CK(cudaMemset(done, 0, 64));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[] = {"local stores only", "+ remote stores", "+ fence.sys per block + ticket", "(fence.gpu per block + ticket)",
                           "+ last block: fence.sys + remote flag"};
    for (int bytes : {4096, 262144}) {
        for (int blocks : {1, 32, 64}) {
            for (int mode = 0; mode <= 4; ++mode) {
                const int reps = 200;
                for (int w = 0; w < 10; ++w) push<<<blocks, 256>>>(src, local, remote, bytes / 16, mode, flag_remote, done, 1);
                CK(cudaDeviceSynchronize());
                cudaEventRecord(e0);
                for (int r = 0; r < reps; ++r) push<<<blocks, 256>>>(src, local, remote, bytes / 16, mode, flag_remote, done, r);
                cudaEventRecord(e1);
                CK(cudaDeviceSynchronize());
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                printf("bytes %7d blocks %2d  %-40s %6.2f us/launch\n", bytes, blocks, names[mode], ms / reps * 1e3);
            }
        }
    }
    return 0;
}
