// simt.hpp -- TEST INFRASTRUCTURE ONLY.  A warp on the CPU: 32 host threads play the lanes, the warp intrinsics the
// device code uses (__ballot_sync, __any_sync, __match_any_sync, __shfl_sync, __shfl_xor_sync, __syncwarp) are rendezvous points between
// them.  With -DTW_SIMT_EMU tw_core.cuh compiles its *device* code paths (work queue with deferred delivery, whole-warp
// region search, reconvergence points, warp-uniform loops) for the host, so tests/emu can run exactly what a GPU warp
// executes against the oracle.  What it cannot show: races between warps and the memory model (warps run one at a time).
#pragma once
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>

namespace simt {

class Barrier {            // sense-reversing, for a fixed number of participants; spins with yield (more threads than cores)
  public:
    explicit Barrier(unsigned n) : n_(n) {}
    void wait()
    {
        const unsigned gen = gen_.load(std::memory_order_acquire);
        if (count_.fetch_add(1, std::memory_order_acq_rel) + 1 == n_) {
            count_.store(0, std::memory_order_relaxed);
            gen_.store(gen + 1, std::memory_order_release);
        } else {
            while (gen_.load(std::memory_order_acquire) == gen) std::this_thread::yield();
        }
    }
  private:
    const unsigned n_;
    std::atomic<unsigned> count_{0}, gen_{0};
};

struct Warp {
    Barrier lanes{32};
    uint64_t slot[32];
};

inline thread_local Warp *g_warp = nullptr;
inline thread_local unsigned g_lane = 0;

}  // namespace simt

struct SimtThreadIdx { unsigned x; };
inline thread_local SimtThreadIdx threadIdx{0};

struct alignas(16) uint4 { unsigned x, y, z, w; };

#define __device__
#define __forceinline__ inline

inline void __syncwarp(unsigned = 0xffffffffu) { simt::g_warp->lanes.wait(); }

inline unsigned __ballot_sync(unsigned, int pred)
{
    simt::Warp &w = *simt::g_warp;
    w.slot[simt::g_lane] = pred ? 1u : 0u;
    w.lanes.wait();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (unsigned)(w.slot[i] & 1u) << i;
    w.lanes.wait();                                  // everybody has read before the slots are reused
    return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }

template <class T>
inline T __shfl_sync(unsigned, T v, int src)
{
    static_assert(sizeof(T) <= 8, "shuffles move at most 64 bits");
    simt::Warp &w = *simt::g_warp;
    w.slot[simt::g_lane] = 0; std::memcpy(&w.slot[simt::g_lane], &v, sizeof v);
    w.lanes.wait();
    T r; std::memcpy(&r, &w.slot[src & 31], sizeof r);
    w.lanes.wait();
    return r;
}
template <class T>
inline T __shfl_xor_sync(unsigned m, T v, int lane_mask) { return __shfl_sync(m, v, (int)(simt::g_lane ^ (unsigned)lane_mask)); }

inline unsigned __match_any_sync(unsigned, unsigned v)
{
    simt::Warp &w = *simt::g_warp;
    w.slot[simt::g_lane] = v;
    w.lanes.wait();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) if ((unsigned)w.slot[i] == v) r |= 1u << i;
    w.lanes.wait();
    return r;
}

inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
