// Does a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization wait for a stream-ordered memcpy / memset that
// sits between it and the previous kernel of the stream?
//     nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/pdl_probe scripts/probes/pdl_memcpy_probe.cu && gpurun_out/pdl_probe
// Result on B200 / driver 580 (gpurun, round 2): 0 of 240 runs saw stale data -- a PDL-launched kernel IS ordered behind a
// preceding cudaMemcpyAsync / cudaMemsetAsync; the hypothesis this probe was written for (an intermittently all-zero first
// batch in tests/test_dp_gpu.py) was wrong, the cause was a copy-stream ordering bug in model._TrainIO.
// Per case: dst is filled with a stale pattern, then [primary kernel ; copy-or-memset of zeros into dst ; reader kernel] are queued
// on ONE stream; the reader (which calls griddepcontrol.wait first, like every PDL-launched kernel of libsemb200) counts the
// elements of dst that are not yet zero.
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

__global__ void primary(float* scratch, int spin) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // like the library's kernels: trigger at the top
    float v = 0.f;
    for (int i = 0; i < spin; ++i) v = v * 1.0001f + 1.f;
    if (v == 123.f) scratch[0] = v;
}
__global__ void reader(const float* buf, size_t n, unsigned long long* bad) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    unsigned long long c = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) c += buf[i] != 0.f;
    if (c) atomicAdd(bad, c);
}
static void launch_reader(bool pdl, cudaStream_t st, const float* buf, size_t n, unsigned long long* bad) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(296); cfg.blockDim = dim3(256); cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, reader, buf, n, bad);
}
int main() {
    const size_t n = 64u << 20;                      // 256 MB: the copy takes ~100 us on the device, ~5 ms over PCIe
    float *src, *dst, *scratch, *hsrc;
    unsigned long long *bad, hbad;
    cudaMalloc(&src, n * 4); cudaMalloc(&dst, n * 4); cudaMalloc(&scratch, 4); cudaMalloc(&bad, 8);
    cudaMallocHost(&hsrc, n * 4);
    memset(hsrc, 0, n * 4);
    cudaMemset(src, 0, n * 4);
    cudaStream_t st; cudaStreamCreate(&st);
    const char* names[] = {"memcpy D2D", "memset", "memcpy H2D (pinned)"};
    for (int mode = 0; mode < 3; ++mode)
        for (int pdl = 0; pdl < 2; ++pdl)
            for (int spin = 0; spin < 2; ++spin) {
                unsigned long long total = 0; int hits = 0;
                for (int rep = 0; rep < 20; ++rep) {
                    cudaMemsetAsync(dst, 0x7f, n * 4, st);                           // stale content
                    cudaMemsetAsync(bad, 0, 8, st);
                    cudaStreamSynchronize(st);
                    primary<<<148, 128, 0, st>>>(scratch, spin ? 200000 : 0);
                    if (mode == 0) cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDeviceToDevice, st);
                    else if (mode == 1) cudaMemsetAsync(dst, 0, n * 4, st);
                    else cudaMemcpyAsync(dst, hsrc, n * 4, cudaMemcpyHostToDevice, st);
                    launch_reader(pdl != 0, st, dst, n, bad);
                    cudaStreamSynchronize(st);
                    cudaMemcpy(&hbad, bad, 8, cudaMemcpyDeviceToHost);
                    total += hbad; hits += hbad != 0;
                }
                printf("%-22s reader %-6s primary %-5s : %2d / 20 runs saw stale data (%llu elements)\n", names[mode], pdl ? "PDL" : "normal",
                       spin ? "long" : "short", hits, total);
            }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
