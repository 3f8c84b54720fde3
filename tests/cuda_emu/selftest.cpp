// Self-test of the CUDA-on-host shim (tests/cuda_emu): block barrier, full-warp and sub-warp (8-lane group) shuffles with
// divergent trip counts, ballot with early-exited lanes, atomics, dynamic shared memory.  Built and run by
// tests/test_kernel_logic_emulated.py; prints "selftest ok".
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

__global__ void k_block_sum(const int* in, int n, int* out) {
    __shared__ int s[256];
    int t = threadIdx.x;
    int v = 0;
    for (int i = blockIdx.x * blockDim.x + t; i < n; i += gridDim.x * blockDim.x) v += in[i];
    s[t] = v;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (t < o) s[t] += s[t + o];
        __syncthreads();
    }
    if (t == 0) atomicAdd(out, s[0]);
}

// 8 lanes per group; group g loops g + 1 times (divergent between the groups of one warp), reducing inside the group each time
__global__ void k_group_reduce(int* out) {
    const unsigned lane = threadIdx.x & 31, group = lane >> 3, gl = lane & 7;
    const unsigned gmask = 0xffu << (8 * group);
    int acc = 0;
    for (unsigned it = 0; it <= group + (threadIdx.x >> 5); ++it) {
        int v = (int)(gl + it);
        for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
        acc += v;
        __syncwarp(gmask);
    }
    if (gl == 0) out[threadIdx.x >> 3] = acc;
}

__global__ void k_ballot_exit(unsigned* out) {
    if (threadIdx.x % 3 == 0) return;  // these lanes are gone
    unsigned m = 0;
    for (unsigned l = 0; l < 32; ++l)
        if (((threadIdx.x & ~31u) + l) % 3 != 0) m |= 1u << l;
    unsigned b = __ballot_sync(m, threadIdx.x & 1);
    if ((threadIdx.x & 31) == 2) out[threadIdx.x >> 5] = b;  // lanes 2 and 34 stay
}

__global__ void k_dyn(int* out) {
    TFBS_DYNAMIC_SHARED(buf);
    int* p = reinterpret_cast<int*>(buf);
    p[threadIdx.x] = threadIdx.x * 2;
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = p[blockDim.x - 1 - threadIdx.x];
}

int main() {
    int bad = 0;
    {
        const int n = 10000;
        std::vector<int> h(n);
        long long want = 0;
        for (int i = 0; i < n; ++i) { h[i] = i % 17 - 5; want += h[i]; }
        int *d, *o;
        cudaMalloc(&d, n * sizeof(int));
        cudaMalloc(&o, sizeof(int));
        cudaMemcpyAsync(d, h.data(), n * sizeof(int), cudaMemcpyHostToDevice, nullptr);
        cudaMemsetAsync(o, 0, sizeof(int), nullptr);
        TFBS_LAUNCH(k_block_sum, 7, 256, 0, nullptr)(d, n, o);
        int got;
        cudaMemcpyAsync(&got, o, sizeof(int), cudaMemcpyDeviceToHost, nullptr);
        if (got != want) { printf("block sum %d != %lld\n", got, want); ++bad; }
    }
    {
        int* o;
        cudaMalloc(&o, 8 * sizeof(int));
        TFBS_LAUNCH(k_group_reduce, 1, 64, 0, nullptr)(o);
        for (int g = 0; g < 8; ++g) {
            int iters = (g & 3) + (g >> 2) + 1, want = 0;
            for (int it = 0; it < iters; ++it) want += 28 + 8 * it;
            if (o[g] != want) { printf("group %d: %d != %d\n", g, o[g], want); ++bad; }
        }
    }
    {
        unsigned* o;
        cudaMalloc(&o, 2 * sizeof(unsigned));
        TFBS_LAUNCH(k_ballot_exit, 1, 64, 0, nullptr)(o);
        for (int w = 0; w < 2; ++w) {
            unsigned want = 0;
            for (int l = 0; l < 32; ++l) {
                int t = 32 * w + l;
                if (t % 3 != 0 && (t & 1)) want |= 1u << l;
            }
            if (o[w] != want) { printf("ballot warp %d: %08x != %08x\n", w, o[w], want); ++bad; }
        }
    }
    {
        int* o;
        cudaMalloc(&o, 3 * 96 * sizeof(int));
        TFBS_LAUNCH(k_dyn, 3, 96, 96 * sizeof(int), nullptr)(o);
        for (int i = 0; i < 3 * 96; ++i)
            if (o[i] != (95 - i % 96) * 2) { printf("dyn %d: %d\n", i, o[i]); ++bad; break; }
    }
    puts(bad ? "selftest FAILED" : "selftest ok");
    return bad ? 1 : 0;
}
