// Host-side helpers shared by the launchers: per-thread LRU plan caches, tuning knobs read once, per-device SM count.
#pragma once
#include <cuda_runtime.h>

namespace vacv {

// ---- tuning knobs ---------------------------------------------------------------------------------------------------
// Diagnostic switches (A/B lines of bench_ops.py, kernel experiments).  Process-wide ints, initialised ONCE from the
// environment (VACV_<NAME>) on first use and changeable at run time through vacv_cuda_set_tuning(); the hot entry points
// only read an int -- no getenv on the call path.  Every change bumps knob_generation(), which is part of every plan key.
enum Knob {
    kKnobPipeNcol = 0,     // VACV_PIPE_NCOL            columns per thread of the fused pipeline (0 = automatic)
    kKnobRpipeNcol,        // VACV_RPIPE_NCOL           columns per thread of the bilinear pipeline (0 = automatic)
    kKnobWarpGather,       // VACV_WARP_GATHER          fp32 warp on the direct gather kernel instead of the TMA-staged one
    kKnobNoRpipe,          // VACV_NO_RPIPE             u8 bilinear off the persistent pipeline
    kKnobRnGather,         // VACV_RESIZE_NORMALIZE_GATHER   resize_normalize off the persistent pipeline
    kKnobWalkSegs,         // VACV_WALK_SEGS            vertical segments per column of the bicubic walkers (0 = automatic)
    kKnobWalkSync,         // VACV_WALK_SYNC            fp32 bicubic walker: register prefetch instead of the cp.async ring
    kKnobWalk2Sync,        // VACV_WALK2_SYNC           u8 bicubic walker: same
    kKnobCubic3Roll,       // VACV_CUBIC3=roll          shared-memory ring kernel instead of the column walker
    kKnobCubicV,           // VACV_CUBIC_V              u8 bicubic kernel variant (0 = default)
    kKnobPipeRows,         // VACV_PIPE_ROWS            fused pipeline on padded surfaces: one bulk copy per row instead of whole bands (padding included)
    kKnobStreamQpt,        // VACV_STREAM_QPT           16-byte groups per thread of the streaming kernels' grids (0 = default)
    kKnobWarpV,            // VACV_WARP_V               u8 BGR warp_affine: 1 = first-generation flat-order gather kernel (0 = automatic: column-owning pack kernel where eligible)
    kKnobLinearV,          // VACV_LINEAR_V             u8 BGR bilinear: 1 = rational scales stay on the persistent pipeline (0 = automatic: periodic walker where eligible)
    kKnobCount
};
int knob(Knob k);
int knob_generation();

// SM count of a device, queried once per device (persistent grids are sized from it, never from a constant).
int sm_count(int device);
inline int current_device() { int d = 0; cudaGetDevice(&d); return d; }
inline int current_sm_count() { return sm_count(current_device()); }

// ---- per-thread LRU of launch plans ------------------------------------------------------------------------------
// A plan (tile geometry, kernel pointer, occupancy, encoded tensor maps ...) costs a cudaFuncSetAttribute + an occupancy
// query to build; callers alternating a few shapes / devices on one thread (two camera resolutions, two GPUs) must not
// rebuild on every call.  find() returns the cached plan for which match(plan) holds, or nullptr; claim() hands out the
// least recently used slot to build a new plan in (mark it valid with commit()).
template <typename Plan, int N = 8>
class PlanCache {
public:
    template <typename Match>
    Plan* find(Match match) {
        for (int i = 0; i < N; ++i)
            if (valid_[i] && match(plan_[i])) { stamp_[i] = ++clock_; return &plan_[i]; }
        return nullptr;
    }
    Plan* claim() {
        int v = 0;
        for (int i = 0; i < N; ++i) {
            if (!valid_[i]) { v = i; break; }
            if (stamp_[i] < stamp_[v]) v = i;
        }
        valid_[v] = false;
        plan_[v] = Plan();
        claimed_ = v;
        return &plan_[v];
    }
    void commit() { valid_[claimed_] = true; stamp_[claimed_] = ++clock_; }
    int builds = 0;   // diagnostic: how many plans this thread has built

private:
    Plan plan_[N] = {};
    bool valid_[N] = {};
    unsigned long long stamp_[N] = {};
    unsigned long long clock_ = 0;
    int claimed_ = 0;
};

}  // namespace vacv
