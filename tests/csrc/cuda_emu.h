// cuda_emu.h -- a minimal CUDA-on-CPU shim for the CPU test suite (test infrastructure, never part of the product).
//
// The container the tests run in by default has no GPU.  The kernels of libtsl that carry non-trivial thread cooperation
// (block reductions, shared-memory tiles, warp shuffles, pivoted LU panels) live in *_kernels.cuh headers that only use the
// CUDA built-ins below; including this header first lets g++ compile them and run them with one std::thread per CUDA thread
// (blocks run one after the other; __syncthreads() is a std::barrier; shuffles exchange through a per-warp buffer).
// Slow, exact in control flow, and enough to find indexing / synchronisation mistakes before spending GPU minutes.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define TSL_CUDA_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __ldg(p) (*(p))
#define __ldcg(p) (*(p))


struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct emu_uint3 { unsigned x, y, z; };

namespace emu {
struct Block {
    std::barrier<> bar;
    std::vector<std::unique_ptr<std::barrier<>>> wbar;
    std::vector<std::array<uint64_t, 32>> xchg;
    Block(int nthreads) : bar(nthreads)
    {
        int nw = (nthreads + 31) / 32;
        for (int w = 0; w < nw; w++) wbar.emplace_back(new std::barrier<>(std::min(32, nthreads - 32 * w)));
        xchg.resize(nw);
    }
};
inline thread_local Block *cur = nullptr;
inline thread_local int lin_tid = 0;
}  // namespace emu

inline thread_local emu_uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

inline void __syncthreads() { emu::cur->bar.arrive_and_wait(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::cur->wbar[emu::lin_tid >> 5]->arrive_and_wait(); }

template <typename T>
inline T emu_shfl(T v, int src_lane)
{
    static_assert(sizeof(T) <= 8, "shuffle payload");
    int w = emu::lin_tid >> 5, lane = emu::lin_tid & 31;
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    emu::cur->xchg[w][lane] = raw;
    emu::cur->wbar[w]->arrive_and_wait();
    uint64_t got = emu::cur->xchg[w][src_lane & 31];
    emu::cur->wbar[w]->arrive_and_wait();
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int o) { return emu_shfl(v, (emu::lin_tid & 31) ^ o); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, int o) { int l = (emu::lin_tid & 31) + o; return emu_shfl(v, l < 32 ? l : (emu::lin_tid & 31)); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, int o) { int l = (emu::lin_tid & 31) - o; return emu_shfl(v, l >= 0 ? l : (emu::lin_tid & 31)); }
template <typename T> inline T __shfl_sync(unsigned, T v, int l) { return emu_shfl(v, l); }

template <typename T> inline T atomicAdd(T *p, T v) { return std::atomic_ref<T>(*p).fetch_add(v); }
inline int atomicOr(int *p, int v) { return std::atomic_ref<int>(*p).fetch_or(v); }
inline unsigned atomicOr(unsigned *p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_or(v); }
inline int atomicMax(int *p, int v) { int o = *p; while (o < v && !std::atomic_ref<int>(*p).compare_exchange_weak(o, v)) {} return o; }

inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
using std::fabs; using std::sqrt; using std::fmin; using std::fmax; using std::floor; using std::acos;
inline float fmaf_emu(float a, float b, float c) { return a * b + c; }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }

// emu_launch(grid, block, kernel, args...): blocks sequentially, one std::thread per CUDA thread of the block
template <class K, class... A>
inline void emu_launch(dim3 grid, dim3 block, K kernel, A... args)
{
    gridDim = grid; blockDim = block;
    int nth = (int)(block.x * block.y * block.z);
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                emu::Block blk(nth);
                std::vector<std::thread> th;
                th.reserve(nth);
                for (int t = 0; t < nth; t++)
                    th.emplace_back([&, t]() {
                        emu::cur = &blk; emu::lin_tid = t;
                        threadIdx.x = t % block.x; threadIdx.y = (t / block.x) % block.y; threadIdx.z = t / (block.x * block.y);
                        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                        kernel(args...);
                        // an exited thread no longer takes part in barriers (as on the hardware)
                        blk.wbar[t >> 5]->arrive_and_drop();
                        blk.bar.arrive_and_drop();
                    });
                for (auto &x : th) x.join();
            }
}
// kernels without __syncthreads / shuffles: plain loops, no threads
template <class K, class... A>
inline void emu_launch_seq(dim3 grid, dim3 block, K kernel, A... args)
{
    gridDim = grid; blockDim = block;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++)
                for (unsigned t = 0; t < block.x * block.y * block.z; t++) {
                    threadIdx.x = t % block.x; threadIdx.y = (t / block.x) % block.y; threadIdx.z = t / (block.x * block.y);
                    blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                    emu::lin_tid = (int)t;
                    kernel(args...);
                }
}
