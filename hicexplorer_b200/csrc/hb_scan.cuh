// Single-block exclusive scan (n up to a few million; runs once per store rebuild).  out has n + 1 entries.
#pragma once
#include "hb_internal.cuh"

template <typename TIn>
__global__ void __launch_bounds__(1024) hb_scan_kernel(const TIn* __restrict__ in, int64_t* __restrict__ out, int64_t n) {
    __shared__ long long part[1024];
    const int t = threadIdx.x;
    const int64_t chunk = (n + 1023) / 1024;
    const int64_t b = t * chunk, e = (b + chunk < n) ? b + chunk : n;
    long long s = 0;
    for (int64_t i = b; i < e; ++i) s += (long long)in[i];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan of the 1024 partials
        long long v = (t >= o) ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    long long run = part[t] - s;
    for (int64_t i = b; i < e; ++i) {
        const long long v = (long long)in[i];
        out[i] = run;
        run += v;
    }
    if (t == 1023) out[n] = part[1023];
}
