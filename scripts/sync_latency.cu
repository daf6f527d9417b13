// Micro-benchmark: cost of getting 64 bytes of reduction results to the host after a kernel, two ways.
//   (a) cudaMemcpyAsync D2H into pinned memory + cudaStreamSynchronize        (what Solver::fetch does)
//   (b) a 1-warp "publish" kernel that stores to mapped pinned memory + a sequence flag the host spins on
// Each round trip = tiny producer kernel -> results on host -> next producer launch.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
__global__ void producer(double *slot, double v) { if (threadIdx.x < 8) slot[threadIdx.x] = v + threadIdx.x; }
__global__ void publish(const double *slot, volatile double *host, volatile unsigned long long *flag, unsigned long long seq) {
    if (threadIdx.x < 8) host[threadIdx.x] = slot[threadIdx.x];
    __threadfence_system();
    __syncwarp();
    if (threadIdx.x == 0) *flag = seq;
}
int main() {
    double *slot, *pinned, *mapped; unsigned long long *flag;
    cudaMalloc(&slot, 64);
    cudaMallocHost(&pinned, 64);
    cudaHostAlloc(&mapped, 64, cudaHostAllocMapped);
    cudaHostAlloc(&flag, 8, cudaHostAllocMapped);
    *flag = 0;
    cudaStream_t s; cudaStreamCreate(&s);
    const int N = 2000;
    for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 2; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        double sum = 0;
        for (int i = 1; i <= N; ++i) {
            producer<<<1, 32, 0, s>>>(slot, (double)i);
            if (mode == 0) {
                cudaMemcpyAsync(pinned, slot, 64, cudaMemcpyDeviceToHost, s);
                cudaStreamSynchronize(s);
                sum += pinned[3];
            } else {
                const unsigned long long seq = (unsigned long long)rep * N + i;
                publish<<<1, 32, 0, s>>>(slot, mapped, flag, seq);
                while (*(volatile unsigned long long *)flag != seq) { }
                sum += ((volatile double *)mapped)[3];
            }
        }
        cudaStreamSynchronize(s);
        double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / N;
        printf("mode %s rep %d: %.2f us per round trip (check %.0f)\n", mode == 0 ? "memcpy+sync" : "publish+spin", rep, us, sum);
    }
    return 0;
}
