// quadsim.cu -- sm_100a kernels + C ABI (include/quadsim.h) of the vectorised quadrotor simulator.
//
// One thread per drone advances the whole control tick in float64 registers (action decode -> S substeps of
// explicit dynamics -> derived rpy/ang_v -> task terms); the CTA then writes its contiguous span of observation
// rows cooperatively so every global access is coalesced:
//   state   : 4 float4 planes [4][N] (16-byte ld/st.global.v4 per thread, fully coalesced)
//   obs     : row-major [N][12+B*A]; a CTA owns rows [c0, c0+T) = one contiguous span; the kinematic head of each
//             row is staged in shared memory by the owning thread, the action history is streamed
//             prev_obs -> obs by whole warps (lane = column), shifted by one action
//   consts  : QsParams travels in the kernel parameter (constant bank, uniform operand -- no load instruction)
// No tensor cores: the path is element-wise; the roofline that bounds it is HBM bandwidth.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <time.h>

#include "quad_core.cuh"

namespace {

thread_local char g_err[256] = "";

int fail(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

int cuda_fail(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kMaxTPB = 128;          // threads (= drones) per CTA upper bound
// fixed part of the step kernel's dynamic shared memory (heads, actions, reductions, in-CTA downwash positions, flags,
// mbarrier), rounded so that the staged rows that follow are 128-byte aligned
__host__ __device__ constexpr size_t smem_fixed(int cap) {
    return ((size_t)cap * 20 * 4 + (size_t)cap * 4 * 4 + (size_t)cap * 2 * 8 + (size_t)cap * 3 * 8 + 3 * (size_t)cap + 16 + 127) / 128 * 128;
}
constexpr size_t kStepSmemFixed = smem_fixed(kMaxTPB);      // upper bound (cap = 128)

struct StepArgs {
    QsParams P;
    QsState st;
    QsStepIO io;
    int act_type, task, n_envs, D, substeps, N, A, obs_dim, tpb, counter_inc;
    unsigned effects, flags;
    int stage_rows;      // 1: the CTA's prev_obs rows are staged in shared memory by one TMA bulk copy
    int cap;             // CTA capacity in drones (64 or 128): sizes the shared-memory arrays
};

__device__ __forceinline__ float4 ldg4(const float* base, long long idx4) {
    return __ldg(reinterpret_cast<const float4*>(base) + idx4);
}
__device__ __forceinline__ void st4(float* base, long long idx4, float4 v) {
    reinterpret_cast<float4*>(base)[idx4] = v;
}

// split a double into float32 hi + float32 lo (hi + lo carries ~48 bits)
__device__ __forceinline__ void split2(double v, float& hi, float& lo) {
    hi = (float)v;
    lo = (float)(v - (double)hi);
}

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier: one thread moves a whole contiguous span global -> shared
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src_gmem) {      // LDGSTS, 4-byte granule
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void load_drone(const float* planes, long long N, long long i, qs::Drone& d) {
    const float4 p0 = ldg4(planes, i), p1 = ldg4(planes, N + i), p2 = ldg4(planes, 2 * N + i), p3 = ldg4(planes, 3 * N + i);
    d.px = p0.x; d.py = p0.y; d.pz = p0.z;
    d.qx = p1.x; d.qy = p1.y; d.qz = p1.z; d.qw = p1.w;
    d.vx = p2.x; d.vy = p2.y; d.vz = p2.z;
    d.wx = (double)p0.w + (double)p3.y;
    d.wy = (double)p2.w + (double)p3.z;
    d.wz = (double)p3.x + (double)p3.w;
}

// normalises the quaternion (the north-star's "quaternion renormalise"; Bullet's own read-back goes through a
// rotation matrix and renormalises too) and stores the 4 planes
__device__ __forceinline__ void store_drone(float* planes, long long N, long long i, qs::Drone& d) {
    const double inv = rsqrt(d.qx * d.qx + d.qy * d.qy + d.qz * d.qz + d.qw * d.qw);
    d.qx *= inv; d.qy *= inv; d.qz *= inv; d.qw *= inv;
    float wxh, wxl, wyh, wyl, wzh, wzl;
    split2(d.wx, wxh, wxl); split2(d.wy, wyh, wyl); split2(d.wz, wzh, wzl);
    st4(planes, i, make_float4((float)d.px, (float)d.py, (float)d.pz, wxh));
    st4(planes, N + i, make_float4((float)d.qx, (float)d.qy, (float)d.qz, (float)d.qw));
    st4(planes, 2 * N + i, make_float4((float)d.vx, (float)d.vy, (float)d.vz, wyh));
    st4(planes, 3 * N + i, make_float4(wzh, wxl, wyl, wzl));
}

__device__ __forceinline__ void init_drone(const QsState& st, long long tbl, qs::Drone& d) {
    const float4 ip = ldg4(st.init_pos, tbl), iq = ldg4(st.init_quat, tbl);
    d.px = ip.x; d.py = ip.y; d.pz = ip.z;
    d.qx = iq.x; d.qy = iq.y; d.qz = iq.z; d.qw = iq.w;
    d.vx = d.vy = d.vz = 0.0;
    d.wx = d.wy = d.wz = 0.0;
}

// ---------------------------------------------------------------------------------------------------------
// Row writer: the CTA's rows [c0, c0+rows) of obs are one contiguous span.  Lane = column (V = float4 when the
// action is 4 wide, so one 18-lane instruction moves a whole 72-float row), warps stride over rows, and U
// independent loads are issued before the first store so the L2 round trip is paid once per U rows.
//   column c <  12/W            : kinematic head staged in shared memory by the owning thread
//   12/W <= c < cols - A/W      : prev_obs column c + A/W   (history shifted by one action)
//   c >= cols - A/W             : this tick's action
// Row modes (autoreset): bit0 keep history unshifted, bit1 mirror the row into final_obs, bit2 zero history in obs.
// ---------------------------------------------------------------------------------------------------------
template <typename V, int W, int U>
__device__ __forceinline__ void write_rows(const StepArgs& a, long long c0, int rows, const float* head_s, const float* act_s,
                                           const unsigned char* mode_s, const float* stage_s) {
    const int cols = a.obs_dim / W, hcols = 12 / W, acols = a.A / W, hist_end = cols - acols;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    // CTA-relative 32-bit offsets (a CTA's span is < 2^31 elements); one 64-bit base per buffer
    const V* prev = reinterpret_cast<const V*>(a.io.obs_prev) + c0 * cols;
    V* out = reinterpret_cast<V*>(a.io.obs) + c0 * cols;
    V* fin = reinterpret_cast<V*>(a.io.final_obs) + c0 * cols;
    const V* head = reinterpret_cast<const V*>(head_s);
    const V* stage = reinterpret_cast<const V*>(stage_s);     // prev rows already in shared memory (TMA) or nullptr
    for (int cb = 0; cb < cols; cb += 32) {                   // column block (one iteration when the row fits 32 lanes)
        const int c = cb + lane;
        const bool col_ok = c < cols;
        for (int r0 = warp; r0 < rows; r0 += nwarps * U) {
            V v[U];
            unsigned char md[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int r = r0 + u * nwarps;
                md[u] = 0x80;                                  // 0x80 = nothing to store
                if (r < rows && col_ok) {
                    const unsigned char mode = mode_s[r];
                    if (c < hcols) {
                        v[u] = head[r * hcols + c];
                        md[u] = 0;
                    } else {
                        md[u] = mode;
                        const bool keep = mode & 1;
                        if (c < hist_end || keep) {
                            const int so = r * cols + c + (keep ? 0 : acols);
                            v[u] = stage ? stage[so] : __ldg(prev + so);
                        }
                        else v[u] = reinterpret_cast<const V*>(act_s + 4 * r)[c - hist_end];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!(md[u] & 0x80)) {
                    const int o = (r0 + u * nwarps) * cols + c;
                    if (md[u] & 2) fin[o] = v[u];
                    if (md[u] & 4) memset(&v[u], 0, sizeof(V));
                    out[o] = v[u];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Fused control tick.  RAW = CtrlAviary semantics (clip raw rpm, [N][20] state vectors out, no task).
// Block = tpb threads, tpb a multiple of D (drones of one aviary never straddle CTAs) when D <= 128.
// ---------------------------------------------------------------------------------------------------------
// PIDACT = the action type runs the embedded DSLPIDControl (PID / VEL / ONE_D_PID): a separate instantiation keeps
// the controller's registers out of the plain RPM kernels.
template <int EFF, bool RAW, bool PIDACT>
__global__ void __launch_bounds__(kMaxTPB, 4) step_kernel(const __grid_constant__ StepArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const QsParams& P = a.P;
    const int tpb = a.tpb, D = a.D, A = a.A;
    const int t = threadIdx.x;
    const long long N = a.N;
    const long long c0 = (long long)blockIdx.x * tpb;          // first drone of this CTA
    const long long i = c0 + t;
    const bool live = (t < tpb) && (i < N);
    const int head = RAW ? 20 : 12;                            // floats staged per row
    // shared layout
    float* head_s = reinterpret_cast<float*>(smem_raw);                  // [tpb][head]
    const int cap = a.cap;
    const size_t fixed = smem_fixed(cap);
    float* act_s = head_s + (size_t)cap * 20;                            // [tpb][4]
    double* red_s = reinterpret_cast<double*>(act_s + (size_t)cap * 4);  // [tpb][2] reward, dist
    double* pos_s = red_s + (size_t)cap * 2;                             // [tpb][3] (in-CTA downwash)
    unsigned char* oob_s = reinterpret_cast<unsigned char*>(pos_s + (size_t)cap * 3);   // [tpb]
    unsigned char* mode_s = oob_s + cap;                                 // [tpb] row mode: 0 shift, 1 keep history, 2 also final_obs
    unsigned char* done_s = mode_s + cap;                                // [tpb] per local env
    unsigned long long* bar_s = reinterpret_cast<unsigned long long*>(smem_raw + fixed - 16);   // mbarrier of the row staging
    float* stage_s = reinterpret_cast<float*>(smem_raw + fixed);                                // [tpb][obs_dim] (+A) when staged

    const long long e = live ? i / D : 0;
    const int le = t / D;                                      // local env (meaningful when D <= tpb)
    const int dslot = (int)(i - e * D);                        // drone index inside its aviary
    const long long tbl = a.st.tables_per_env ? i : dslot;

    // Programmatic dependent launch: when the host launched this grid with programmatic stream serialization its CTAs
    // may already be resident while the previous kernel in the stream drains; nothing written by that kernel is read
    // before this point (no-op for ordinary launches).
    asm volatile("griddepcontrol.wait;" ::: "memory");

    qs::Drone d;
    qs::Derived o;
    qs::PidState pst = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    float act[4] = {0.f, 0.f, 0.f, 0.f};
    double rpm[4] = {0, 0, 0, 0}, rpm_prev[4] = {0, 0, 0, 0};
    double R_last[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    int sc = 0;
    bool pending = false;
    constexpr bool pid_act = PIDACT;

    // The observation rows of this CTA are one contiguous span, and the new rows are the old ones shifted left by one
    // action: new_flat[j] = old_flat[j + A] once every thread has patched its own row (head -> slots [A, A+12), new
    // action -> the A slots after the row, i.e. the dead head slots of the next row).  So: ONE thread starts a TMA bulk
    // copy of the old span into shared memory now (completion on an mbarrier), the physics below runs while it is in
    // flight, then the span is streamed out with a flat, fully coalesced float4 copy.  Spans too large for shared
    // memory are only pulled into L2 (and written by write_rows).
    const bool want_rows = !RAW && a.io.obs && a.io.obs_prev && a.obs_dim > 12 && !(a.flags & QS_FLAG_SKIP_EPILOGUE);
    const int rows = (int)((N - c0) < tpb ? (N - c0) : tpb);
    if (want_rows && a.stage_rows == 1) {
        if (t == 0) mbar_init(bar_s, 1);
        __syncthreads();
    }

    if (live) {
        load_drone(a.st.planes, N, i, d);
        if (A == 4) {
            const float4 v = ldg4(a.io.action, i);
            act[0] = v.x; act[1] = v.y; act[2] = v.z; act[3] = v.w;
        } else if (A == 3) {
            act[0] = __ldg(a.io.action + i * 3); act[1] = __ldg(a.io.action + i * 3 + 1); act[2] = __ldg(a.io.action + i * 3 + 2);
        } else {
            act[0] = __ldg(a.io.action + i);
        }
        if (((EFF & QS_EFFECT_DRAG) || (a.flags & QS_FLAG_RPM_FROM_LAST)) && a.st.last_rpm) {
            const float4 v = ldg4(a.st.last_rpm, i);
            rpm_prev[0] = v.x; rpm_prev[1] = v.y; rpm_prev[2] = v.z; rpm_prev[3] = v.w;
        }
        if (pid_act) {
            const float* ps = a.st.pid;
            pst.ipx = ps[i]; pst.ipy = ps[N + i]; pst.ipz = ps[2 * N + i];
            pst.lr = ps[3 * N + i]; pst.lp = ps[4 * N + i]; pst.ly = ps[5 * N + i];
            pst.irx = ps[6 * N + i]; pst.iry = ps[7 * N + i]; pst.irz = ps[8 * N + i];
        }
        sc = a.st.step_counter[e];
        if ((a.flags & QS_FLAG_AUTORESET_NEXT_STEP) && a.st.pending_reset) pending = a.st.pending_reset[e] != 0;
    }

    // async copy of the old observation span, issued AFTER this thread's state/action loads so that the small, latency
    // critical loads are ahead of the 36 KB bulk transfer in the memory system
    if (want_rows && a.stage_rows == 1) {
        if (t == 0) tma_bulk_g2s(stage_s, a.io.obs_prev + c0 * a.obs_dim, (unsigned)(rows * a.obs_dim * 4), bar_s);
    } else if (want_rows && a.stage_rows == 2) {
        // spans that are not 16-byte aligned (odd action widths): per-thread 4-byte async copies (LDGSTS), still fire-and-forget
        const float* src = a.io.obs_prev + c0 * a.obs_dim;
        for (int j = t; j < rows * a.obs_dim; j += blockDim.x) cp_async4(stage_s + j, src + j);
    } else if (want_rows && t == 0) {
        const uintptr_t p0 = reinterpret_cast<uintptr_t>(a.io.obs_prev + c0 * a.obs_dim);
        const uintptr_t beg = (p0 + 15) & ~(uintptr_t)15;
        const uintptr_t end = (p0 + (uintptr_t)rows * a.obs_dim * 4) & ~(uintptr_t)15;
        if (end > beg) {
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(beg), "r"((unsigned)(end - beg)) : "memory");
        }
    }

    if (live && !pending) {
        double cur_yaw = 0.0;
        if (a.act_type == QS_ACT_VEL) {
            double r_, p_;
            qs::quat_to_euler<false>(d.qx, d.qy, d.qz, d.qw, r_, p_, cur_yaw);
        }
        if (a.flags & QS_FLAG_RPM_FROM_LAST) {
            rpm[0] = rpm_prev[0]; rpm[1] = rpm_prev[1]; rpm[2] = rpm_prev[2]; rpm[3] = rpm_prev[3];
        } else {
            qs::decode_action<PIDACT>(P, a.act_type, act, d, cur_yaw, pst, rpm);
        }
    }

    // ---- physics: S substeps ------------------------------------------------------------------------------
    if ((EFF & QS_EFFECT_DW) && a.io.dw_fz == nullptr) {
        // downwash inside the CTA: all drones of an aviary sit in this CTA (D <= tpb), positions go through smem
        for (int s = 0; s < a.substeps; ++s) {
            if (live) { pos_s[3 * t] = d.px; pos_s[3 * t + 1] = d.py; pos_s[3 * t + 2] = d.pz; }
            __syncthreads();
            double fz = 0.0;
            if (live && !pending) {
                const int b = le * D;
                for (int k = 0; k < D; ++k) {                                   // BaseAviary.py:798-811
                    const double dz = pos_s[3 * (b + k) + 2] - d.pz;
                    const double dx = pos_s[3 * (b + k)] - d.px, dy = pos_s[3 * (b + k) + 1] - d.py;
                    const double dxy2 = dx * dx + dy * dy;
                    if (dz > 0.0 && dxy2 < 100.0) fz += qs::downwash_pair(P, dz, dxy2);
                }
                qs::dyn_tick<EFF>(P, d, rpm, s == 0 ? rpm_prev : rpm, fz, 1, R_last);
            }
            __syncthreads();
        }
    } else if (live && !pending) {
        const double fz = (EFF & QS_EFFECT_DW) ? (double)__ldg(a.io.dw_fz + i) : 0.0;
        qs::dyn_tick<EFF>(P, d, rpm, rpm_prev, fz, a.substeps, R_last);
    }

    // ---- derived outputs, task terms --------------------------------------------------------------------
    bool env_done = false;
    if (live) {
        if (pending) {                       // NEXT_STEP autoreset: this call only resets the env
            init_drone(a.st, tbl, d);
            if (a.flags & QS_FLAG_AUTORESET_CLEARS_PID) pst = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        }
        if (a.flags & QS_FLAG_RPY_F32) qs::derive<true>(d, R_last, o); else qs::derive<false>(d, R_last, o);
        if (pending) { o.ax = o.ay = o.az = 0.0; }
    }
    const bool want_epilogue = !(a.flags & QS_FLAG_SKIP_EPILOGUE);
    if (!RAW && a.task == QS_TASK_HOVER && want_epilogue) {
        if (live) {
            const float4 tp = ldg4(a.st.target_pos, tbl);
            const qs::TaskTerms tt = qs::hover_terms(P, d, o, tp.x, tp.y, tp.z);
            red_s[2 * t] = tt.reward; red_s[2 * t + 1] = tt.dist; oob_s[t] = tt.out_of_bounds ? 1 : 0;
        }
        __syncthreads();
        if (live && dslot == 0) {
            double rew = 0.0, dist = 0.0; bool oob = false;
            for (int k = 0; k < D; ++k) { rew += red_s[2 * (t + k)]; dist += red_s[2 * (t + k) + 1]; oob |= oob_s[t + k] != 0; }
            bool term = dist < P.term_dist;                                        // HoverAviary.py:91
            bool trunc = oob || ((double)sc / P.pyb_freq > P.episode_len_sec);     // HoverAviary.py:113
            if (pending) { rew = 0.0; term = false; trunc = false; }
            a.io.reward[e] = (float)rew;
            a.io.terminated[e] = term ? 1 : 0;
            a.io.truncated[e] = trunc ? 1 : 0;
            if (a.io.done) a.io.done[e] = (term || trunc) ? 1 : 0;
            done_s[le] = (term || trunc) ? 1 : 0;
        }
        __syncthreads();
        if (live) env_done = done_s[le] != 0;
    } else if (!RAW && want_epilogue && live && dslot == 0) {
        a.io.reward[e] = -1.0f; a.io.terminated[e] = 0; a.io.truncated[e] = 0;     // CtrlAviary-style dummy task
        if (a.io.done) a.io.done[e] = 0;
    }

    // ---- stage this drone's row head, autoreset, store state ---------------------------------------------
    if (live) {
        float* h = head_s + (size_t)t * head;
        // row mode bits: 1 = keep history unshifted, 2 = also copy the row to final_obs, 4 = zero the history in obs
        unsigned char mode = pending ? (unsigned char)(1 | ((a.flags & QS_FLAG_AUTORESET_CLEARS_HISTORY) ? 4 : 0)) : (unsigned char)0;
        if (RAW) {
            // _getDroneStateVector (BaseAviary.py:541-561); quaternion reported normalised
            const double inv = rsqrt(d.qx * d.qx + d.qy * d.qy + d.qz * d.qz + d.qw * d.qw);
            h[0] = (float)d.px; h[1] = (float)d.py; h[2] = (float)d.pz;
            h[3] = (float)(d.qx * inv); h[4] = (float)(d.qy * inv); h[5] = (float)(d.qz * inv); h[6] = (float)(d.qw * inv);
            h[7] = (float)o.roll; h[8] = (float)o.pitch; h[9] = (float)o.yaw;
            h[10] = (float)d.vx; h[11] = (float)d.vy; h[12] = (float)d.vz;
            h[13] = (float)o.ax; h[14] = (float)o.ay; h[15] = (float)o.az;
            h[16] = (float)rpm[0]; h[17] = (float)rpm[1]; h[18] = (float)rpm[2]; h[19] = (float)rpm[3];
        } else {
            const bool same_step = (a.flags & QS_FLAG_AUTORESET_SAME_STEP) && env_done;
            if (same_step) {
                // terminal observation head goes straight to final_obs (rare path, strided store is fine)
                if (a.io.final_obs) {
                    float* f = a.io.final_obs + i * a.obs_dim;
                    f[0] = (float)d.px; f[1] = (float)d.py; f[2] = (float)d.pz;
                    f[3] = (float)o.roll; f[4] = (float)o.pitch; f[5] = (float)o.yaw;
                    f[6] = (float)d.vx; f[7] = (float)d.vy; f[8] = (float)d.vz;
                    f[9] = (float)o.ax; f[10] = (float)o.ay; f[11] = (float)o.az;
                    mode |= 2;
                }
                if (a.flags & QS_FLAG_AUTORESET_CLEARS_HISTORY) mode |= 4;
                if (a.flags & QS_FLAG_AUTORESET_CLEARS_PID) pst = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                init_drone(a.st, tbl, d);                                          // BaseAviary.py:451-505
                double Rr[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
                if (a.flags & QS_FLAG_RPY_F32) qs::derive<true>(d, Rr, o); else qs::derive<false>(d, Rr, o);
                rpm[0] = rpm[1] = rpm[2] = rpm[3] = 0.0;                           // last_clipped_action = 0
                sc = -a.counter_inc;                                               // -> 0 after the increment below
            }
            // KIN observation head: pos3 rpy3 vel3 ang_v3 (BaseRLAviary.py:310-315)
            h[0] = (float)d.px; h[1] = (float)d.py; h[2] = (float)d.pz;
            h[3] = (float)o.roll; h[4] = (float)o.pitch; h[5] = (float)o.yaw;
            h[6] = (float)d.vx; h[7] = (float)d.vy; h[8] = (float)d.vz;
            h[9] = (float)o.ax; h[10] = (float)o.ay; h[11] = (float)o.az;
            act_s[4 * t] = act[0]; act_s[4 * t + 1] = act[1]; act_s[4 * t + 2] = act[2]; act_s[4 * t + 3] = act[3];
        }
        mode_s[t] = mode;
        store_drone(a.st.planes, N, i, d);
        if (a.st.last_rpm && !pending)
            st4(a.st.last_rpm, i, make_float4((float)rpm[0], (float)rpm[1], (float)rpm[2], (float)rpm[3]));
        if (pid_act) {
            float* ps = a.st.pid;
            ps[i] = (float)pst.ipx; ps[N + i] = (float)pst.ipy; ps[2 * N + i] = (float)pst.ipz;
            ps[3 * N + i] = (float)pst.lr; ps[4 * N + i] = (float)pst.lp; ps[5 * N + i] = (float)pst.ly;
            ps[6 * N + i] = (float)pst.irx; ps[7 * N + i] = (float)pst.iry; ps[8 * N + i] = (float)pst.irz;
        }
        if (dslot == 0 && want_epilogue) {
            if (pending) {
                a.st.step_counter[e] = 0;
                a.st.pending_reset[e] = 0;
            } else {
                a.st.step_counter[e] = sc + a.counter_inc;                         // BaseAviary.py:382
                if ((a.flags & QS_FLAG_AUTORESET_NEXT_STEP) && a.st.pending_reset && env_done) a.st.pending_reset[e] = 1;
            }
        }
    }
    // all the FP64 work of this CTA is done: let the next grid in the stream start moving in behind the stores
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (a.io.obs == nullptr || !want_epilogue) return;
    const int od = a.obs_dim;
    if (want_rows && a.stage_rows) {
        // ---- staged rows: wait for the async copy, patch my row in shared memory, stream the span out ----------
        if (a.stage_rows == 1) mbar_wait(bar_s, 0); else cp_async_commit_wait_all();
        __syncthreads();
        if (live) {
            float* row = stage_s + (size_t)t * od;
            const float* h = head_s + (size_t)t * 12;
            const unsigned char mode = mode_s[t];
            if (mode & 1) {                                   // NEXT_STEP reset tick: history is NOT shifted (or is cleared)
                for (int k = od - 1; k >= 12; --k) row[k + A] = (mode & 4) ? 0.f : row[k];
            } else if (A == 4) {
                *reinterpret_cast<float4*>(row + od) = make_float4(act[0], act[1], act[2], act[3]);
            } else if (A == 3) {
                row[od] = act[0]; row[od + 1] = act[1]; row[od + 2] = act[2];
            } else {
                row[od] = act[0];
            }
            if (A == 4) {
                float4* r4 = reinterpret_cast<float4*>(row + 4);
                r4[0] = make_float4(h[0], h[1], h[2], h[3]); r4[1] = make_float4(h[4], h[5], h[6], h[7]); r4[2] = make_float4(h[8], h[9], h[10], h[11]);
            } else {
                for (int k = 0; k < 12; ++k) row[A + k] = h[k];
            }
        }
        __syncthreads();
        const bool clear_hist = a.flags & QS_FLAG_AUTORESET_CLEARS_HISTORY;
        const int lane = t & 31;
        if (A == 4) {
            const float4* src = reinterpret_cast<const float4*>(stage_s) + 1;
            float4* out = reinterpret_cast<float4*>(a.io.obs + c0 * od);
            const int c4n = od >> 2, n4 = rows * c4n;
            if (a.stage_rows == 1) {
                // TMA bulk store: one thread hands the whole patched span (shared memory, shifted by one action) to the
                // copy engine; the other threads go on to the terminal-observation rows.  The async proxy must see the
                // generic-proxy patches (fence), and shared memory must stay alive until it has been read (wait_group.read).
                if (t == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 ::"l"(out), "r"(smem_u32(src)), "r"((unsigned)(n4 * 16)) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else {
                const int nt = blockDim.x;
                int j = t;
                for (; j + 5 * nt < n4; j += 6 * nt) {        // 6 independent LDS.128 in flight, then 6 coalesced STG.128
                    const float4 v0 = src[j], v1 = src[j + nt], v2 = src[j + 2 * nt], v3 = src[j + 3 * nt], v4 = src[j + 4 * nt], v5 = src[j + 5 * nt];
                    out[j] = v0; out[j + nt] = v1; out[j + 2 * nt] = v2; out[j + 3 * nt] = v3; out[j + 4 * nt] = v4; out[j + 5 * nt] = v5;
                }
                for (; j < n4; j += nt) out[j] = src[j];
            }
            // SAME_STEP autoreset: the history part of the terminal observation of finished rows.  Each warp ballots the
            // flags of its own 32 rows and copies only the flagged ones, a whole row per instruction.
            if (a.io.final_obs && (a.flags & QS_FLAG_AUTORESET_SAME_STEP)) {
                float4* fin = reinterpret_cast<float4*>(a.io.final_obs + c0 * od);
                unsigned m = __ballot_sync(0xffffffffu, live && (mode_s[t] & 2));
                const int r0 = t & ~31;
                for (; m; m &= m - 1) {
                    const int r = r0 + __ffs(m) - 1;
                    for (int c = 3 + lane; c < c4n; c += 32) fin[r * c4n + c] = src[r * c4n + c];
                }
            }
        } else {
            const float* src = stage_s + A;
            float* out = a.io.obs + c0 * od;
            for (int j = t; j < rows * od; j += blockDim.x) out[j] = src[j];
            if (a.io.final_obs && (a.flags & QS_FLAG_AUTORESET_SAME_STEP)) {
                float* fin = a.io.final_obs + c0 * od;
                unsigned m = __ballot_sync(0xffffffffu, live && (mode_s[t] & 2));
                const int r0 = t & ~31;
                for (; m; m &= m - 1) {
                    const int r = r0 + __ffs(m) - 1;
                    for (int c = 12 + lane; c < od; c += 32) fin[r * od + c] = src[r * od + c];
                }
            }
        }
        if (A == 4 && a.stage_rows == 1 && t == 0) {
            if (clear_hist) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // stores complete (ordering vs the zeroing below)
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");             // shared memory has been read
        }
        if (clear_hist) {                                     // optional: the observation after a reset carries an empty action buffer
            __syncthreads();
            if (live && (mode_s[t] & 4) && !(mode_s[t] & 1)) {
                float* orow = a.io.obs + i * od;
                for (int k = 12; k < od; ++k) orow[k] = 0.f;
            }
        }
        return;
    }
    __syncthreads();

    // ---- cooperative, coalesced write of this CTA's observation rows (unstaged paths) -----------------------------
    if (RAW) {
        float* out = a.io.obs + c0 * 20;
        for (int j = t; j < rows * 20; j += blockDim.x) out[j] = head_s[j];
    } else if (A == 4) {
        write_rows<float4, 4, 8>(a, c0, rows, head_s, act_s, mode_s, nullptr);
    } else {
        write_rows<float, 1, 8>(a, c0, rows, head_s, act_s, mode_s, nullptr);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Multi-tick rollout: the fused control tick in a loop.  Drone state lives in registers, the CTA's observation rows in a
// shared-memory window that slides by one action per tick (new_flat[j] = old_flat[j + A], so "shifting the history" is
// `base += A`); per tick the kernel reads the action and writes the rows, reward and flags.  Bit-identical to T calls of
// qs_step: the state is rounded to its float32 plane representation at every tick boundary exactly like store/load.
// ---------------------------------------------------------------------------------------------------------
struct RolloutArgs {
    QsParams P;
    QsState st;
    QsRolloutIO io;
    int act_type, task, n_envs, D, substeps, N, A, obs_dim, tpb;
    unsigned effects, flags;
    int stage_mode, cap;
};

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float u32_to_pm1(unsigned u) { return (float)(u >> 8) * (1.0f / 8388608.0f) - 1.0f; }   // [-1, 1)

// state -> float32 planes -> state, without the memory round trip (same roundings as store_drone + load_drone)
__device__ __forceinline__ void round_to_planes(qs::Drone& d) {
    const double inv = rsqrt(d.qx * d.qx + d.qy * d.qy + d.qz * d.qz + d.qw * d.qw);
    d.qx = (double)(float)(d.qx * inv); d.qy = (double)(float)(d.qy * inv); d.qz = (double)(float)(d.qz * inv); d.qw = (double)(float)(d.qw * inv);
    d.px = (double)(float)d.px; d.py = (double)(float)d.py; d.pz = (double)(float)d.pz;
    d.vx = (double)(float)d.vx; d.vy = (double)(float)d.vy; d.vz = (double)(float)d.vz;
    float h, l;
    split2(d.wx, h, l); d.wx = (double)h + (double)l;
    split2(d.wy, h, l); d.wy = (double)h + (double)l;
    split2(d.wz, h, l); d.wz = (double)h + (double)l;
}

template <int EFF, bool PIDACT>
__global__ void __launch_bounds__(kMaxTPB, 4) rollout_kernel(const __grid_constant__ RolloutArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const QsParams& P = a.P;
    const int tpb = a.tpb, D = a.D, A = a.A, od = a.obs_dim, T = a.io.T;
    const int t = threadIdx.x;
    const long long N = a.N, E = a.n_envs;
    const long long c0 = (long long)blockIdx.x * tpb;
    const long long i = c0 + t;
    const bool live = (t < tpb) && (i < N);
    const int rows = (int)((N - c0) < tpb ? (N - c0) : tpb);
    double* red_s = reinterpret_cast<double*>(smem_raw);                               // [tpb][2]
    const int cap = a.cap;
    const size_t fixed = smem_fixed(cap);
    double* pos_s = red_s + (size_t)cap * 2;                                           // [tpb][3]
    unsigned char* oob_s = reinterpret_cast<unsigned char*>(pos_s + (size_t)cap * 3);
    unsigned char* done_s = oob_s + cap;
    unsigned long long* bar_s = reinterpret_cast<unsigned long long*>(smem_raw + fixed - 16);
    float* stage_s = reinterpret_cast<float*>(smem_raw + fixed);                       // [tpb*od + (T+1)*A] sliding window

    const long long e = live ? i / D : 0;
    const int le = t / D;
    const int dslot = (int)(i - e * D);
    const long long tbl = a.st.tables_per_env ? i : dslot;

    if (a.stage_mode == 1) {
        if (t == 0) mbar_init(bar_s, 1);
        __syncthreads();
    }
    qs::Drone d;
    qs::PidState pst = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double rpm_prev[4] = {0, 0, 0, 0};
    int sc = 0;
    if (live) {
        load_drone(a.st.planes, N, i, d);
        if ((EFF & QS_EFFECT_DRAG) && a.st.last_rpm) {
            const float4 v = ldg4(a.st.last_rpm, i);
            rpm_prev[0] = v.x; rpm_prev[1] = v.y; rpm_prev[2] = v.z; rpm_prev[3] = v.w;
        }
        if (PIDACT) {
            const float* ps = a.st.pid;
            pst.ipx = ps[i]; pst.ipy = ps[N + i]; pst.ipz = ps[2 * N + i];
            pst.lr = ps[3 * N + i]; pst.lp = ps[4 * N + i]; pst.ly = ps[5 * N + i];
            pst.irx = ps[6 * N + i]; pst.iry = ps[7 * N + i]; pst.irz = ps[8 * N + i];
        }
        sc = a.st.step_counter[e];
    }
    if (a.stage_mode == 1) {
        if (t == 0) tma_bulk_g2s(stage_s, a.io.obs_init + c0 * od, (unsigned)(rows * od * 4), bar_s);
        mbar_wait(bar_s, 0);
    } else {
        const float* src = a.io.obs_init + c0 * od;
        for (int j = t; j < rows * od; j += blockDim.x) cp_async4(stage_s + j, src + j);
        cp_async_commit_wait_all();
    }
    __syncthreads();

    float* base = stage_s;                     // window start: rows of the observation BEFORE the current tick
    double rpm[4] = {0, 0, 0, 0};
    for (int k = 0; k < T; ++k) {
        // ---- this tick's action: caller-provided or generated on the device --------------------------------------
        float act[4] = {0.f, 0.f, 0.f, 0.f};
        if (live) {
            if (a.io.actions) {
                const float* ap = a.io.actions + ((long long)k * N + i) * A;
                if (A == 4) { const float4 v = __ldg(reinterpret_cast<const float4*>(ap)); act[0] = v.x; act[1] = v.y; act[2] = v.z; act[3] = v.w; }
                else if (A == 3) { act[0] = __ldg(ap); act[1] = __ldg(ap + 1); act[2] = __ldg(ap + 2); }
                else act[0] = __ldg(ap);
            } else {
                const unsigned long long key = a.io.seed + 2ull * (unsigned long long)((a.io.tick0 + k) * N + i);
                const unsigned long long r0 = splitmix64(key), r1 = splitmix64(key + 1);
                act[0] = u32_to_pm1((unsigned)r0); act[1] = u32_to_pm1((unsigned)(r0 >> 32));
                act[2] = u32_to_pm1((unsigned)r1); act[3] = u32_to_pm1((unsigned)(r1 >> 32));
                if (A < 4) act[3] = 0.f;
                if (A < 3) { act[1] = 0.f; act[2] = 0.f; }
            }
            float* tail = base + (size_t)t * od + od;        // new action -> the A slots after my row (dead head of the next row)
            if (A == 4) *reinterpret_cast<float4*>(tail) = make_float4(act[0], act[1], act[2], act[3]);
            else if (A == 3) { tail[0] = act[0]; tail[1] = act[1]; tail[2] = act[2]; }
            else tail[0] = act[0];
            if (a.io.actions_out) {
                float* ao = a.io.actions_out + ((long long)k * N + i) * A;
                if (A == 4) *reinterpret_cast<float4*>(ao) = make_float4(act[0], act[1], act[2], act[3]);
                else if (A == 3) { ao[0] = act[0]; ao[1] = act[1]; ao[2] = act[2]; }
                else ao[0] = act[0];
            }
        }
        // ---- physics ---------------------------------------------------------------------------------------------
        double R_last[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (live) {
            double cur_yaw = 0.0;
            if (a.act_type == QS_ACT_VEL) { double r_, p_; qs::quat_to_euler<false>(d.qx, d.qy, d.qz, d.qw, r_, p_, cur_yaw); }
            qs::decode_action<PIDACT>(P, a.act_type, act, d, cur_yaw, pst, rpm);
        }
        if (EFF & QS_EFFECT_DW) {
            for (int s = 0; s < a.substeps; ++s) {
                if (live) { pos_s[3 * t] = d.px; pos_s[3 * t + 1] = d.py; pos_s[3 * t + 2] = d.pz; }
                __syncthreads();
                if (live) {
                    double fz = 0.0;
                    const int b = le * D;
                    for (int q = 0; q < D; ++q) {
                        const double dz = pos_s[3 * (b + q) + 2] - d.pz;
                        const double dx = pos_s[3 * (b + q)] - d.px, dy = pos_s[3 * (b + q) + 1] - d.py;
                        const double dxy2 = dx * dx + dy * dy;
                        if (dz > 0.0 && dxy2 < 100.0) fz += qs::downwash_pair(P, dz, dxy2);
                    }
                    qs::dyn_tick<EFF>(P, d, rpm, s == 0 ? rpm_prev : rpm, fz, 1, R_last);
                }
                __syncthreads();
            }
        } else if (live) {
            qs::dyn_tick<EFF>(P, d, rpm, rpm_prev, 0.0, a.substeps, R_last);
        }
        qs::Derived o;
        if (live) { if (a.flags & QS_FLAG_RPY_F32) qs::derive<true>(d, R_last, o); else qs::derive<false>(d, R_last, o); }
        // ---- task ------------------------------------------------------------------------------------------------
        bool env_done = false;
        if (a.task == QS_TASK_HOVER) {
            if (live) {
                const float4 tp = ldg4(a.st.target_pos, tbl);
                const qs::TaskTerms tt = qs::hover_terms(P, d, o, tp.x, tp.y, tp.z);
                red_s[2 * t] = tt.reward; red_s[2 * t + 1] = tt.dist; oob_s[t] = tt.out_of_bounds ? 1 : 0;
            }
            __syncthreads();
            if (live && dslot == 0) {
                double rew = 0.0, dist = 0.0; bool oob = false;
                for (int q = 0; q < D; ++q) { rew += red_s[2 * (t + q)]; dist += red_s[2 * (t + q) + 1]; oob |= oob_s[t + q] != 0; }
                const bool term = dist < P.term_dist;
                const bool trunc = oob || ((double)sc / P.pyb_freq > P.episode_len_sec);
                const long long oe = (long long)k * E + e;
                a.io.reward[oe] = (float)rew; a.io.terminated[oe] = term ? 1 : 0; a.io.truncated[oe] = trunc ? 1 : 0;
                if (a.io.done) a.io.done[oe] = (term || trunc) ? 1 : 0;
                done_s[le] = (term || trunc) ? 1 : 0;
            }
            __syncthreads();
            if (live) env_done = done_s[le] != 0;
        } else if (live && dslot == 0) {
            const long long oe = (long long)k * E + e;
            a.io.reward[oe] = -1.0f; a.io.terminated[oe] = 0; a.io.truncated[oe] = 0;
            if (a.io.done) a.io.done[oe] = 0;
        }
        // ---- autoreset, head, bookkeeping ----------------------------------------------------------------------------
        if (live) {
            float* row = base + A + (size_t)t * od;              // my row in the NEXT window
            if ((a.flags & QS_FLAG_AUTORESET_SAME_STEP) && env_done) {
                if (a.flags & QS_FLAG_AUTORESET_CLEARS_HISTORY) for (int q = 12; q < od; ++q) row[q] = 0.f;
                if (a.flags & QS_FLAG_AUTORESET_CLEARS_PID) pst = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                init_drone(a.st, tbl, d);
                const double Rr[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
                if (a.flags & QS_FLAG_RPY_F32) qs::derive<true>(d, Rr, o); else qs::derive<false>(d, Rr, o);
                rpm[0] = rpm[1] = rpm[2] = rpm[3] = 0.0;
                sc = -a.substeps;
            }
            row[0] = (float)d.px; row[1] = (float)d.py; row[2] = (float)d.pz;
            row[3] = (float)o.roll; row[4] = (float)o.pitch; row[5] = (float)o.yaw;
            row[6] = (float)d.vx; row[7] = (float)d.vy; row[8] = (float)d.vz;
            row[9] = (float)o.ax; row[10] = (float)o.ay; row[11] = (float)o.az;
            sc += a.substeps;
            if (k < T - 1) round_to_planes(d);                // (the final store_drone applies the same rounding once)
            if (PIDACT) {      // the PID state is stored as float32 between ticks as well
                pst.ipx = (double)(float)pst.ipx; pst.ipy = (double)(float)pst.ipy; pst.ipz = (double)(float)pst.ipz;
                pst.lr = (double)(float)pst.lr; pst.lp = (double)(float)pst.lp; pst.ly = (double)(float)pst.ly;
                pst.irx = (double)(float)pst.irx; pst.iry = (double)(float)pst.iry; pst.irz = (double)(float)pst.irz;
            }
            rpm_prev[0] = (double)(float)rpm[0]; rpm_prev[1] = (double)(float)rpm[1];
            rpm_prev[2] = (double)(float)rpm[2]; rpm_prev[3] = (double)(float)rpm[3];
        }
        __syncthreads();
        // ---- stream the CTA's rows out: obs[k][c0 .. c0+rows) = window shifted by one action ----------------------------
        base += A;
        {
            float* outp = a.io.obs + ((long long)k * N + c0) * od;
            float* lastp = (k == T - 1 && a.io.obs_last) ? a.io.obs_last + c0 * od : nullptr;
            if (A == 4 && a.stage_mode == 1) {
                // TMA bulk store of the window (see step_kernel); the window is rewritten next tick, so wait until the
                // copy engine has read it
                if (t == 0) {
                    const unsigned bytes = (unsigned)(rows * od * 4);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(outp), "r"(smem_u32(base)), "r"(bytes) : "memory");
                    if (lastp) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(lastp), "r"(smem_u32(base)), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
            } else if (A == 4) {
                const float4* src = reinterpret_cast<const float4*>(base);
                float4* out = reinterpret_cast<float4*>(outp);
                float4* last = reinterpret_cast<float4*>(lastp);
                const int n4 = rows * (od >> 2), nt = blockDim.x;
                int j = t;
                for (; j + 5 * nt < n4; j += 6 * nt) {
                    const float4 v0 = src[j], v1 = src[j + nt], v2 = src[j + 2 * nt], v3 = src[j + 3 * nt], v4 = src[j + 4 * nt], v5 = src[j + 5 * nt];
                    out[j] = v0; out[j + nt] = v1; out[j + 2 * nt] = v2; out[j + 3 * nt] = v3; out[j + 4 * nt] = v4; out[j + 5 * nt] = v5;
                    if (last) { last[j] = v0; last[j + nt] = v1; last[j + 2 * nt] = v2; last[j + 3 * nt] = v3; last[j + 4 * nt] = v4; last[j + 5 * nt] = v5; }
                }
                for (; j < n4; j += nt) { const float4 v = src[j]; out[j] = v; if (last) last[j] = v; }
            } else {
                for (int j = t; j < rows * od; j += blockDim.x) { const float v = base[j]; outp[j] = v; if (lastp) lastp[j] = v; }
            }
        }
        __syncthreads();
    }
    if (live) {
        store_drone(a.st.planes, N, i, d);
        if (a.st.last_rpm) st4(a.st.last_rpm, i, make_float4((float)rpm[0], (float)rpm[1], (float)rpm[2], (float)rpm[3]));
        if (PIDACT) {
            float* ps = a.st.pid;
            ps[i] = (float)pst.ipx; ps[N + i] = (float)pst.ipy; ps[2 * N + i] = (float)pst.ipz;
            ps[3 * N + i] = (float)pst.lr; ps[4 * N + i] = (float)pst.lp; ps[5 * N + i] = (float)pst.ly;
            ps[6 * N + i] = (float)pst.irx; ps[7 * N + i] = (float)pst.iry; ps[8 * N + i] = (float)pst.irz;
        }
        if (dslot == 0) a.st.step_counter[e] = sc;
    }
}

// ---------------------------------------------------------------------------------------------------------
// DSLPIDControl.computeControl for n drones (stand-alone entry).
// ---------------------------------------------------------------------------------------------------------
struct PidArgs {
    QsParams P;
    float* pid;
    double dt;
    const float *pos, *quat, *vel, *tpos, *trpy, *tvel, *trr;
    int pos_stride, quat_stride, vel_stride, n;
    float *rpm_out, *pos_e_out, *yaw_e_out;
};

__global__ void __launch_bounds__(128) pid_kernel(const __grid_constant__ PidArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long N = a.n;
    if (i >= N) return;
    const float* p = a.pos + i * a.pos_stride;
    const float* q = a.quat + i * a.quat_stride;
    const float* v = a.vel + i * a.vel_stride;
    qs::PidState st;
    st.ipx = a.pid[i]; st.ipy = a.pid[N + i]; st.ipz = a.pid[2 * N + i];
    st.lr = a.pid[3 * N + i]; st.lp = a.pid[4 * N + i]; st.ly = a.pid[5 * N + i];
    st.irx = a.pid[6 * N + i]; st.iry = a.pid[7 * N + i]; st.irz = a.pid[8 * N + i];
    const double tyaw = a.trpy ? (double)a.trpy[3 * i + 2] : 0.0;
    double tv[3] = {0, 0, 0}, tr[3] = {0, 0, 0};
    if (a.tvel) { tv[0] = a.tvel[3 * i]; tv[1] = a.tvel[3 * i + 1]; tv[2] = a.tvel[3 * i + 2]; }
    if (a.trr) { tr[0] = a.trr[3 * i]; tr[1] = a.trr[3 * i + 1]; tr[2] = a.trr[3 * i + 2]; }
    double rpm[4], pe[3], ye;
    qs::pid_control(a.P, st, a.dt, p[0], p[1], p[2], q[0], q[1], q[2], q[3], v[0], v[1], v[2],
                    a.tpos[3 * i], a.tpos[3 * i + 1], a.tpos[3 * i + 2], tyaw, tv[0], tv[1], tv[2], tr[0], tr[1], tr[2],
                    rpm, pe, ye);
    a.pid[i] = (float)st.ipx; a.pid[N + i] = (float)st.ipy; a.pid[2 * N + i] = (float)st.ipz;
    a.pid[3 * N + i] = (float)st.lr; a.pid[4 * N + i] = (float)st.lp; a.pid[5 * N + i] = (float)st.ly;
    a.pid[6 * N + i] = (float)st.irx; a.pid[7 * N + i] = (float)st.iry; a.pid[8 * N + i] = (float)st.irz;
    reinterpret_cast<float4*>(a.rpm_out)[i] = make_float4((float)rpm[0], (float)rpm[1], (float)rpm[2], (float)rpm[3]);
    if (a.pos_e_out) { a.pos_e_out[3 * i] = (float)pe[0]; a.pos_e_out[3 * i + 1] = (float)pe[1]; a.pos_e_out[3 * i + 2] = (float)pe[2]; }
    if (a.yaw_e_out) a.yaw_e_out[i] = (float)ye;
}

// ---------------------------------------------------------------------------------------------------------
// Pairwise downwash for large aviaries: tiled all-pairs with exact bounding-box culling.  A CTA owns 128 drones
// ("rows") of ONE aviary and runs 1024 threads: thread (slice s, row n) evaluates the tile entries k = s (mod 8), so a
// 16 384-drone formation fills 128 SMs with 32 warps each instead of 64 SMs with 8.  The sources stream through shared
// memory in tiles of 1024 = 32 chunks of 32 (next tile prefetched into registers); every chunk carries its bounding box (warp redux on order-preserving
// integer keys), every warp knows the box of its 32 rows, and a warp skips a chunk when NO pair of the two boxes can
// contribute: all dz <= 0, or every dxy^2 >= 100 (BaseAviary.py:800), or every pair takes dw_pair's underflow early-out
// (dxy^2 > 220 beta_max^2: skipping changes no bit of the result).  Index-coherent
// formations (grids, Morton order) therefore cost O(N k) instead of O(N^2).  The pair term is float32 (predicate
// first, expf only for pairs in range), partial sums are float64 and are combined in a fixed order (deterministic).
// (qs_downwash: no workspace.  qs_downwash_boxed / qs_downwash_rows below use a precomputed box table instead of tiles.)
// ---------------------------------------------------------------------------------------------------------
struct DwArgs {
    float prop_radius, dw1, dw2, dw3;
    const float* rows;
    float* fz;
    int D, tiles_per_env, cull;
};

constexpr int kDwDrones = 128, kDwSlices = 8, kDwTile = 1024, kDwChunks = kDwTile / 32;
constexpr long long kDwSpinLimit = 4000000000LL;   // ~2 s of SM clock

// One pair of BaseAviary._downwash (BaseAviary.py:798-806) in float32: alpha exp(-.5 (dxy/beta)^2), 0 when the pair fails
// the reference's predicate (dz > 0, dxy < 10) or when the Gaussian is below float32 range anyway (dxy^2 > 220 beta^2:
// exp2(-158) flushes to 0, so the early-out changes no bit).  Reciprocals and exp2 are the SFU approximations (1-2 ulp):
// the kernel is SFU/ALU bound, and the oracle tolerance (1e-5) is three orders above that.
__device__ __forceinline__ float dw_pair(float prop_radius, float dw1, float dw2, float dw3, float dz, float dxy2) {
    const float beta = dw2 * dz + dw3;
    const float b2 = beta * beta;
    if (!(dz > 0.f && dxy2 < 100.f) || dxy2 > 220.f * b2) return 0.f;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(dz * b2));                  // one SFU reciprocal for 1/dz and 1/beta^2
    const float rr = (0.25f * prop_radius) * (b2 * r);
    const float u2 = dxy2 * (dz * r);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-0.72134752f * u2));      // exp(-u2/2) = 2^(-u2 log2(e)/2)
    return dw1 * (rr * rr) * e;
}

__device__ __forceinline__ int f2key(float x) { const int i = __float_as_int(x); return i ^ ((i >> 31) & 0x7fffffff); }
__device__ __forceinline__ float key2f(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }
__device__ __forceinline__ float warp_min(float x) { return key2f(__reduce_min_sync(0xffffffffu, f2key(x))); }
__device__ __forceinline__ float warp_max(float x) { return key2f(__reduce_max_sync(0xffffffffu, f2key(x))); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(kDwDrones * kDwSlices) downwash_kernel(const __grid_constant__ DwArgs a) {
    __shared__ float4 tile[kDwTile];
    __shared__ float cbox[kDwChunks][6];                  // xmin xmax ymin ymax zmin zmax per chunk
    __shared__ double part[kDwSlices][kDwDrones];
    constexpr float BIG = 3e30f;
    const int env = blockIdx.x / a.tiles_per_env;
    const int tb = blockIdx.x - env * a.tiles_per_env;
    const long long base = (long long)env * a.D;
    const int ln = threadIdx.x % kDwDrones, sl = threadIdx.x / kDwDrones;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = tb * kDwDrones + ln;                    // my row inside the aviary
    const bool live = n < a.D;
    float4 me = make_float4(0.f, 0.f, BIG, 0.f);
    if (live) me = ldg4(a.rows, base + n);
    // box of this warp's 32 rows (warp-uniform)
    const float rx0 = warp_min(live ? me.x : BIG), rx1 = warp_max(live ? me.x : -BIG);
    const float ry0 = warp_min(live ? me.y : BIG), ry1 = warp_max(live ? me.y : -BIG);
    const float rz0 = warp_min(live ? me.z : BIG), rz1 = warp_max(live ? me.z : -BIG);
    const float* src = a.rows + base * 4;
    const int n_src = a.D;
    double acc = 0.0;
    const float4 dead = make_float4(0.f, 0.f, -BIG, 0.f);
    auto load_src = [&](int j) { return j < n_src ? ldg4(src, j) : dead; };
    float4 nxt = load_src(threadIdx.x);
    for (int j0 = 0; j0 < n_src; j0 += kDwTile) {
        {
            const float4 o = nxt;
            const bool ok = j0 + (int)threadIdx.x < n_src;
            tile[threadIdx.x] = o;
            const float x0 = warp_min(ok ? o.x : BIG), x1 = warp_max(ok ? o.x : -BIG);
            const float y0 = warp_min(ok ? o.y : BIG), y1 = warp_max(ok ? o.y : -BIG);
            const float z0 = warp_min(ok ? o.z : BIG), z1 = warp_max(ok ? o.z : -BIG);
            if (lane == 0) { float* c = cbox[warp]; c[0] = x0; c[1] = x1; c[2] = y0; c[3] = y1; c[4] = z0; c[5] = z1; }
        }
        __syncthreads();
        if (j0 + kDwTile < n_src) nxt = load_src(j0 + kDwTile + threadIdx.x);        // in flight during the evaluation
        bool act;
        {
            const float* c = cbox[lane];
            const float dzhi = c[5] - rz0, dzlo = fmaxf(c[4] - rz1, 0.f);
            const float gx = fmaxf(fmaxf(c[0] - rx1, rx0 - c[1]), 0.f), gy = fmaxf(fmaxf(c[2] - ry1, ry0 - c[3]), 0.f);
            const float g2 = gx * gx + gy * gy;
            const float b0 = fabsf(a.dw2 * dzlo + a.dw3), b1 = fabsf(a.dw2 * dzhi + a.dw3);
            const float bm = fmaxf(b0, b1);
            act = (dzhi > 0.f) && !(g2 > 100.001f) && !(g2 > 220.f * bm * bm);
            if (!a.cull) act = true;
        }
        unsigned m = __ballot_sync(0xffffffffu, act);
        float part_f = 0.f;
        while (m) {
            const int c = __ffs(m) - 1;
            m &= m - 1;
#pragma unroll
            for (int q = 0; q < 32 / kDwSlices; ++q) {
                const float4 o = tile[c * 32 + sl + kDwSlices * q];
                const float dz = o.z - me.z;
                const float dx = o.x - me.x, dy = o.y - me.y;
                const float dxy2 = dx * dx + dy * dy;
                part_f -= dw_pair(a.prop_radius, a.dw1, a.dw2, a.dw3, dz, dxy2);
            }
        }
        acc += (double)part_f;
        __syncthreads();
    }
    part[sl][ln] = acc;
    __syncthreads();
    if (sl == 0 && live) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < kDwSlices; ++k) t += part[k][ln];
        a.fz[base + n] = (float)t;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Boxed downwash (formations): the sources carry a table of bounding boxes, one per chunk of 32 consecutive positions
// ({min x,y,z,-}{max x,y,z,-}).  A CTA = 32 rows x 8 slices (256 threads, no position tile in shared memory): per batch
// of 256 chunks each warp tests 32 boxes against the box of the 32 rows (same exact predicate as above), the eight
// ballot words go through shared memory, then every warp walks the active chunks in index order and evaluates its 4
// entries of each (uniform 16-byte loads).  Cost: O(N/32) box tests + the pairs that can contribute, per row group;
// the small CTAs (N/32 of them) keep every SM busy when a formation is split over several GPUs.
// ---------------------------------------------------------------------------------------------------------
struct DwbArgs {
    float prop_radius, dw1, dw2, dw3;
    const float* rows;
    float* fz;
    const float* src;          // nullptr: sources of an aviary are its own rows
    const float* boxes;        // [chunks][8], per aviary when src == nullptr
    int n_src, D, groups_per_env, chunks, cull;
    const unsigned* ready;
    unsigned seq;
    int world;
    unsigned* err;
};

__global__ void __launch_bounds__(256, 4) downwash_boxed_kernel(const __grid_constant__ DwbArgs a) {
    __shared__ unsigned masks[8];
    __shared__ double part[8][32];
    constexpr float BIG = 3e30f;
    const int env = blockIdx.x / a.groups_per_env;
    const int g = blockIdx.x - env * a.groups_per_env;
    const long long base = (long long)env * a.D;
    const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int n = g * 32 + lane;
    const bool live = n < a.D;
    if (a.ready) {
        if ((int)threadIdx.x < a.world) {
            const long long t0 = clock64();
            while ((int)(ld_acquire_sys(a.ready + threadIdx.x) - a.seq) < 0) {
                if (clock64() - t0 > kDwSpinLimit) { if (a.err) atomicExch(a.err, 1u); break; }
            }
        }
        __syncthreads();
    }
    float4 me = make_float4(0.f, 0.f, BIG, 0.f);
    if (live) me = ldg4(a.rows, base + n);
    const float rx0 = warp_min(live ? me.x : BIG), rx1 = warp_max(live ? me.x : -BIG);
    const float ry0 = warp_min(live ? me.y : BIG), ry1 = warp_max(live ? me.y : -BIG);
    const float rz0 = warp_min(live ? me.z : BIG), rz1 = warp_max(live ? me.z : -BIG);
    const bool shared_src = a.src != nullptr;            // exchange buffers are written by peers: L2-coherent loads
    const float4* src = reinterpret_cast<const float4*>(shared_src ? a.src : a.rows + base * 4);
    const float4* boxes = reinterpret_cast<const float4*>(a.boxes) + (shared_src ? 0 : (long long)env * a.chunks * 2);
    const int n_src = shared_src ? a.n_src : a.D;
    const float4 dead = make_float4(0.f, 0.f, -BIG, 0.f);
    double acc = 0.0;
    for (int cb = 0; cb < a.chunks; cb += 256) {
        {
            const int c = cb + sl * 32 + lane;
            bool act = false;
            if (c < a.chunks) {
                const float4 lo = shared_src ? __ldcg(boxes + 2 * c) : __ldg(boxes + 2 * c);
                const float4 hi = shared_src ? __ldcg(boxes + 2 * c + 1) : __ldg(boxes + 2 * c + 1);
                const float dzhi = hi.z - rz0, dzlo = fmaxf(lo.z - rz1, 0.f);
                const float gx = fmaxf(fmaxf(lo.x - rx1, rx0 - hi.x), 0.f), gy = fmaxf(fmaxf(lo.y - ry1, ry0 - hi.y), 0.f);
                const float g2 = gx * gx + gy * gy;
                const float b0 = fabsf(a.dw2 * dzlo + a.dw3), b1 = fabsf(a.dw2 * dzhi + a.dw3);
                const float bm = fmaxf(b0, b1);
                act = (dzhi > 0.f) && !(g2 > 100.001f) && !(g2 > 220.f * bm * bm);
                if (!a.cull) act = true;
            }
            const unsigned m = __ballot_sync(0xffffffffu, act);
            if (lane == 0) masks[sl] = m;
        }
        __syncthreads();
#pragma unroll 1
        for (int w = 0; w < 8; ++w) {
            unsigned m = masks[w];
            float part_f = 0.f;
            while (m) {
                const int c = cb + w * 32 + __ffs(m) - 1;
                m &= m - 1;
                float4 o[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = c * 32 + sl + 8 * q;
                    o[q] = dead;
                    if (j < n_src) o[q] = shared_src ? __ldcg(src + j) : __ldg(src + j);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float dz = o[q].z - me.z;
                    const float dx = o[q].x - me.x, dy = o[q].y - me.y;
                    const float dxy2 = dx * dx + dy * dy;
                    part_f -= dw_pair(a.prop_radius, a.dw1, a.dw2, a.dw3, dz, dxy2);
                }
            }
            acc += (double)part_f;
        }
        __syncthreads();
    }
    part[sl][lane] = acc;
    __syncthreads();
    if (sl == 0 && live) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][lane];
        a.fz[base + n] = (float)t;
    }
}

// boxes of the chunks of 32 consecutive positions, per aviary: one warp per chunk
__global__ void __launch_bounds__(256) dw_boxes_kernel(const float* __restrict__ pos, float* __restrict__ boxes, int D, int chunks, long long total_chunks) {
    constexpr float BIG = 3e30f;
    const long long wc = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wc >= total_chunks) return;
    const int lane = threadIdx.x & 31;
    const long long env = wc / chunks;
    const int c = (int)(wc - env * chunks);
    const int i = c * 32 + lane;
    const bool ok = i < D;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) v = ldg4(pos, env * D + i);
    const float x0 = warp_min(ok ? v.x : BIG), x1 = warp_max(ok ? v.x : -BIG);
    const float y0 = warp_min(ok ? v.y : BIG), y1 = warp_max(ok ? v.y : -BIG);
    const float z0 = warp_min(ok ? v.z : BIG), z1 = warp_max(ok ? v.z : -BIG);
    if (lane == 0) {
        float4* b = reinterpret_cast<float4*>(boxes) + 2 * wc;
        b[0] = make_float4(x0, y0, z0, 0.f);
        b[1] = make_float4(x1, y1, z1, 0.f);
    }
}

// Push this GPU's slice of a formation's positions AND the boxes of its chunks into every rank's gathered array (own +
// NVLink peers), then raise this rank's sequence flag on every rank: remote stores are fire-and-forget, the last CTA to
// finish (fence + counter) publishes the flags with release semantics.  offset is a multiple of 32, so a warp = a chunk.
struct PubArgs {
    const float* pos;
    float* dst[QS_MAX_PEERS];
    unsigned* flags[QS_MAX_PEERS];
    unsigned* counter;
    int n, offset, n_total, world, rank;
    unsigned seq;
};

__global__ void __launch_bounds__(128) dw_publish_kernel(const __grid_constant__ PubArgs a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // the consumer synchronises on the flags
    constexpr float BIG = 3e30f;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool ok = i < a.n;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) v = ldg4(a.pos, i);
    const float x0 = warp_min(ok ? v.x : BIG), x1 = warp_max(ok ? v.x : -BIG);
    const float y0 = warp_min(ok ? v.y : BIG), y1 = warp_max(ok ? v.y : -BIG);
    const float z0 = warp_min(ok ? v.z : BIG), z1 = warp_max(ok ? v.z : -BIG);
    const int first = i - lane;                                           // warp-uniform
    if (first < a.n) {
        const long long chunk = (a.offset + first) >> 5;
        for (int r = 0; r < a.world; ++r) {
            float4* d = reinterpret_cast<float4*>(a.dst[r]);
            if (ok) d[a.offset + i] = v;
            if (lane == 0) {
                float4* b = d + a.n_total + 2 * chunk;
                b[0] = make_float4(x0, y0, z0, 0.f);
                b[1] = make_float4(x1, y1, z1, 0.f);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();                                           // cumulative over the CTA's stores (barrier above)
        const unsigned t = atomicAdd(a.counter, 1u);
        if (t == gridDim.x - 1) {
            *a.counter = 0u;
            __threadfence_system();
            for (int r = 0; r < a.world; ++r) st_release_sys(a.flags[r] + a.rank, a.seq);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Neighbourhood query (BaseAviary._getAdjacencyMatrix, BaseAviary.py:658-675): out[e][i][j] = (i == j) or
// |pos_i - pos_j| < radius.  HBM-write bound (D^2 bytes per aviary): a thread produces 16 columns of one row as one
// 16-byte store, a warp 512 contiguous bytes; the 512 column positions of a CTA sit in shared memory.  The
// comparison is made in float32 and re-evaluated in float64 (sqrt(dx^2+dy^2+dz^2) < radius, the reference's
// arithmetic) only when the float32 value is within 2e-4 relative of the threshold (warp-uniform rare branch).
// ---------------------------------------------------------------------------------------------------------
struct AdjArgs {
    const float* planes;
    unsigned char* out;
    double radius;
    int D, col_tiles, row_tiles;
};

constexpr int kAdjCols = 512, kAdjRows = 64;

__global__ void __launch_bounds__(256) adjacency_kernel(const __grid_constant__ AdjArgs a) {
    __shared__ float4 cols[kAdjCols];
    int b = blockIdx.x;
    const int ct = b % a.col_tiles; b /= a.col_tiles;
    const int rt = b % a.row_tiles;
    const int env = b / a.row_tiles;
    const long long base = (long long)env * a.D;
    const int c0 = ct * kAdjCols, r0 = rt * kAdjRows;
    for (int k = threadIdx.x; k < kAdjCols; k += blockDim.x)
        cols[k] = (c0 + k < a.D) ? ldg4(a.planes, base + c0 + k) : make_float4(3e30f, 3e30f, 3e30f, 0.f);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float r2f = (float)(a.radius * a.radius);
    float r2lo = r2f * (1.f - 2e-4f), r2hi = r2f * (1.f + 2e-4f);           // outside [lo, hi] float32 decides
    if (a.radius < 0.0) r2lo = r2hi = -1.f;                                 // |d| < negative radius: never
    const bool vec = (a.D % 16) == 0;
    for (int rr = warp; rr < kAdjRows; rr += 8) {
        const int i = r0 + rr;
        if (i >= a.D) break;
        const float4 me = ldg4(a.planes, base + i);
        // group q: lane evaluates column c0 + 32 q + lane (conflict-free LDS.128), the ballot collects the 32 results;
        // lane q keeps the word, so lane L finds its 16 output columns in lane L/2's word, half L%2
        unsigned word = 0u;
#pragma unroll
        for (int q = 0; q < kAdjCols / 32; ++q) {
            const float4 o = cols[q * 32 + lane];
            const float dx = o.x - me.x, dy = o.y - me.y, dz = o.z - me.z;
            const float d2 = dx * dx + dy * dy + dz * dz;
            bool near = d2 < r2lo;
            const bool amb = !near && !(d2 > r2hi);
            if (__any_sync(0xffffffffu, amb)) {                          // rare: within 2e-4 of the threshold -> reference arithmetic
                if (amb) {
                    const double ex = (double)me.x - (double)o.x, ey = (double)me.y - (double)o.y, ez = (double)me.z - (double)o.z;
                    near = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez))) < a.radius;
                }
            }
            const unsigned bits = __ballot_sync(0xffffffffu, near);
            if (lane == q) word = bits;
        }
        const int ci = i - c0;                                            // identity (BaseAviary.py:666)
        if (ci >= 0 && ci < kAdjCols && lane == (ci >> 5)) word |= 1u << (ci & 31);
        const unsigned h = (__shfl_sync(0xffffffffu, word, lane >> 1) >> ((lane & 1) * 16)) & 0xffffu;
        unsigned w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = (((h >> (4 * k)) & 0xfu) * 0x00204081u) & 0x01010101u;
        const int j = c0 + lane * 16;
        unsigned char* dst = a.out + ((size_t)(base + i)) * a.D + j;
        if (vec) {
            if (j < a.D) *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            for (int q = 0; q < 16 && j + q < a.D; ++q) dst[q] = (unsigned char)((w[q >> 2] >> (8 * (q & 3))) & 0xffu);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Masked reset (BaseAviary.reset / _housekeeping).
// ---------------------------------------------------------------------------------------------------------
struct ResetArgs {
    QsState st;
    const unsigned char* mask;
    int D, reset_pid, obs_dim, raw20;
    long long N;
    float* obs;
};

__global__ void __launch_bounds__(128) reset_kernel(const __grid_constant__ ResetArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    const long long e = i / a.D;
    if (a.mask && !a.mask[e]) return;
    const int dslot = (int)(i - e * a.D);
    const long long tbl = a.st.tables_per_env ? i : dslot;
    qs::Drone d;
    init_drone(a.st, tbl, d);
    qs::Derived o;
    const double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    qs::derive<false>(d, R, o);
    store_drone(a.st.planes, a.N, i, d);
    if (a.st.last_rpm) st4(a.st.last_rpm, i, make_float4(0.f, 0.f, 0.f, 0.f));
    if (a.reset_pid && a.st.pid)
        for (int k = 0; k < 9; ++k) a.st.pid[k * a.N + i] = 0.f;
    if (dslot == 0) {
        a.st.step_counter[e] = 0;
        if (a.st.pending_reset) a.st.pending_reset[e] = 0;
    }
    if (a.obs) {
        if (a.raw20) {
            float* h = a.obs + i * 20;
            h[0] = (float)d.px; h[1] = (float)d.py; h[2] = (float)d.pz;
            h[3] = (float)d.qx; h[4] = (float)d.qy; h[5] = (float)d.qz; h[6] = (float)d.qw;
            h[7] = (float)o.roll; h[8] = (float)o.pitch; h[9] = (float)o.yaw;
            for (int k = 10; k < 20; ++k) h[k] = 0.f;
        } else {
            float* h = a.obs + i * a.obs_dim;
            h[0] = (float)d.px; h[1] = (float)d.py; h[2] = (float)d.pz;
            h[3] = (float)o.roll; h[4] = (float)o.pitch; h[5] = (float)o.yaw;
            for (int k = 6; k < 12; ++k) h[k] = 0.f;
        }
    }
}

constexpr size_t kStageLimit = 40 * 1024;      // bytes of staged rows per CTA (4 CTAs/SM must fit in 227 KB)

// rows of the k finished aviaries -> compact buffer (one aviary = D*obs_dim contiguous floats)
__global__ void gather_rows_kernel(const float* __restrict__ src, const long long* __restrict__ idx, float* __restrict__ dst, int k, int row_floats) {
    const int r = blockIdx.x;
    if (r >= k) return;
    const float* s = src + idx[r] * (long long)row_floats;
    float* d = dst + (long long)r * row_floats;
    for (int j = threadIdx.x; j < row_floats; j += blockDim.x) d[j] = s[j];
}

size_t step_smem_bytes(const StepArgs& a) {
    return smem_fixed(a.cap) + (a.stage_rows ? (size_t)a.tpb * a.obs_dim * 4 + 32 : 0);
}

template <bool RAW, bool PIDACT>
cudaError_t launch_step(const StepArgs& a, cudaStream_t s) {
    const int blocks = (int)((a.N + a.tpb - 1) / a.tpb);
    const int threads = ((a.tpb + 31) / 32) * 32;
    const size_t sm = step_smem_bytes(a);
    static const bool pdl = !(getenv("QS_PDL") && atoi(getenv("QS_PDL")) == 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = sm; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
#define QS_CASE(E)                                                                                               \
    case E: {                                                                                                    \
        static bool attr_set = false;                                                                            \
        if (!attr_set) {                                                                                         \
            cudaFuncSetAttribute(step_kernel<E, RAW, PIDACT>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                 (int)(kStepSmemFixed + kStageLimit + 32));                                           \
            attr_set = true;                                                                                     \
        }                                                                                                        \
        return cudaLaunchKernelEx(&cfg, step_kernel<E, RAW, PIDACT>, a);                                         \
    }
    switch (a.effects & 7u) {
        QS_CASE(0) QS_CASE(1) QS_CASE(2) QS_CASE(3) QS_CASE(4) QS_CASE(5) QS_CASE(6) QS_CASE(7)
    }
#undef QS_CASE
    return cudaGetLastError();
}

int act_width(int act_type) {
    switch (act_type) {
        case QS_ACT_RPM: case QS_ACT_VEL: case QS_ACT_RAW_RPM: return 4;
        case QS_ACT_PID: return 3;
        case QS_ACT_ONE_D_RPM: case QS_ACT_ONE_D_PID: return 1;
        default: return -1;
    }
}

int check_state(const QsState* st, int need_tables) {
    if (!st || !st->planes || !st->step_counter) return fail(QS_ERR_NULL, "QsState: planes/step_counter is NULL");
    if (!aligned16(st->planes)) return fail(QS_ERR_ALIGN, "QsState.planes must be 16-byte aligned");
    if (st->last_rpm && !aligned16(st->last_rpm)) return fail(QS_ERR_ALIGN, "QsState.last_rpm must be 16-byte aligned");
    if (need_tables) {
        if (!st->init_pos || !st->init_quat) return fail(QS_ERR_NULL, "QsState: init_pos/init_quat is NULL");
        if (!aligned16(st->init_pos) || !aligned16(st->init_quat)) return fail(QS_ERR_ALIGN, "init tables must be 16-byte aligned");
    }
    return 0;
}

// CTA capacity in drones.  Measured on B200 (tools/ab.py, same box): 32-drone CTAs (one warp) beat 64 and 128 for the
// single-tick kernels at every size (65 536 drones: 16.2 / 16.1 / 17.3 us, 1 M drones: 161 / 162 / 169 us) -- more,
// smaller CTAs per SM sit at different phases (load / FP64 / store) at any instant, and 65 536 drones spread 14/13 per
// SM instead of 4/3.  The multi-tick rollout keeps its window in shared memory for many ticks and prefers 64.
int cta_capacity(long long N, int D, bool rollout = false) {
    static const int forced = getenv("QS_CTA_CAP") ? atoi(getenv("QS_CTA_CAP")) : 0;      // experiments only
    if (forced == 32 || forced == 64 || forced == 128) return D <= forced ? forced : kMaxTPB;
    (void)N;
    if (D <= 32 && !rollout) return 32;
    return D <= 64 ? 64 : kMaxTPB;
}

int block_size_for(int D, int cap = kMaxTPB) { return D <= cap ? D * (cap / D) : cap; }

}  // namespace

// =============================================================================================================
extern "C" {

int qs_abi_version(void) { return QS_ABI_VERSION; }
const char* qs_last_error(void) { return g_err; }
int qs_sizeof_params(void) { return (int)sizeof(QsParams); }
int qs_sizeof_state(void) { return (int)sizeof(QsState); }
int qs_sizeof_step_io(void) { return (int)sizeof(QsStepIO); }

int qs_step(const QsParams* p, const QsState* st, const QsStepIO* io, int act_type, int task,
            int n_envs, int drones_per_env, int substeps, unsigned effects, unsigned flags, void* stream) {
    if (!p || !io) return fail(QS_ERR_NULL, "qs_step: NULL params/io");
    const bool autoreset = flags & (QS_FLAG_AUTORESET_SAME_STEP | QS_FLAG_AUTORESET_NEXT_STEP);
    if (int rc = check_state(st, autoreset ? 1 : 0)) return rc;
    if (n_envs <= 0 || drones_per_env <= 0 || substeps <= 0) return fail(QS_ERR_SIZE, "qs_step: n_envs, drones_per_env, substeps must be > 0");
    const int A = act_width(act_type);
    const bool state20 = flags & QS_FLAG_OBS_STATE20;
    if (A < 0 || (act_type == QS_ACT_RAW_RPM && !state20)) return fail(QS_ERR_ENUM, "qs_step: bad act_type (use qs_dyn_substeps for raw rpm)");
    if (state20 && (task != QS_TASK_NONE || autoreset)) return fail(QS_ERR_UNSUPPORTED, "qs_step: OBS_STATE20 needs QS_TASK_NONE and no autoreset");
    if (task != QS_TASK_NONE && task != QS_TASK_HOVER) return fail(QS_ERR_ENUM, "qs_step: bad task");
    if (effects & ~7u) return fail(QS_ERR_ENUM, "qs_step: bad effects");
    if ((flags & QS_FLAG_AUTORESET_SAME_STEP) && (flags & QS_FLAG_AUTORESET_NEXT_STEP)) return fail(QS_ERR_ENUM, "qs_step: two autoreset modes");
    if (!io->action) return fail(QS_ERR_NULL, "qs_step: action is NULL");
    if (A == 4 && !aligned16(io->action)) return fail(QS_ERR_ALIGN, "qs_step: [N][4] action must be 16-byte aligned");
    const bool skip = flags & QS_FLAG_SKIP_EPILOGUE;
    if ((flags & QS_FLAG_RPM_FROM_LAST) && !st->last_rpm) return fail(QS_ERR_NULL, "qs_step: RPM_FROM_LAST needs QsState.last_rpm");
    if (!skip && !state20 && (!io->reward || !io->terminated || !io->truncated)) return fail(QS_ERR_NULL, "qs_step: reward/terminated/truncated is NULL");
    if (io->act_buffer_size < 0) return fail(QS_ERR_SIZE, "qs_step: act_buffer_size < 0");
    if (io->obs && io->act_buffer_size > 0 && !state20 && !io->obs_prev) return fail(QS_ERR_NULL, "qs_step: obs_prev is NULL");
    if (io->obs && io->obs == io->obs_prev) return fail(QS_ERR_UNSUPPORTED, "qs_step: obs and obs_prev must be distinct buffers");
    if (task == QS_TASK_HOVER && !st->target_pos) return fail(QS_ERR_NULL, "qs_step: target_pos is NULL");
    if (task == QS_TASK_HOVER && !aligned16(st->target_pos)) return fail(QS_ERR_ALIGN, "qs_step: target_pos must be 16-byte aligned");
    if (task == QS_TASK_HOVER && drones_per_env > kMaxTPB) return fail(QS_ERR_UNSUPPORTED, "qs_step: task reduction supports drones_per_env <= 128");
    const bool pid_act = act_type == QS_ACT_PID || act_type == QS_ACT_VEL || act_type == QS_ACT_ONE_D_PID;
    if (pid_act && !st->pid) return fail(QS_ERR_NULL, "qs_step: PID action type needs QsState.pid");
    if ((effects & QS_EFFECT_DRAG) && !st->last_rpm) return fail(QS_ERR_NULL, "qs_step: DRAG needs QsState.last_rpm");
    if ((effects & QS_EFFECT_DW) && !io->dw_fz && drones_per_env > kMaxTPB)
        return fail(QS_ERR_UNSUPPORTED, "qs_step: in-CTA downwash needs drones_per_env <= 128 (else pass dw_fz from qs_downwash, substeps = 1)");
    if ((effects & QS_EFFECT_DW) && io->dw_fz && substeps != 1) return fail(QS_ERR_UNSUPPORTED, "qs_step: external dw_fz requires substeps == 1");
    if ((flags & QS_FLAG_AUTORESET_NEXT_STEP) && !st->pending_reset) return fail(QS_ERR_NULL, "qs_step: NEXT_STEP autoreset needs pending_reset");
    StepArgs a;
    memset(&a, 0, sizeof(a));
    a.P = *p; a.st = *st; a.io = *io;
    a.act_type = act_type; a.task = task; a.n_envs = n_envs; a.D = drones_per_env; a.substeps = substeps;
    a.N = n_envs * drones_per_env; a.A = A; a.obs_dim = state20 ? 20 : 12 + io->act_buffer_size * A;
    a.cap = cta_capacity(a.N, drones_per_env);
    a.tpb = block_size_for(drones_per_env, a.cap);
    a.counter_inc = io->tick_substeps > 0 ? io->tick_substeps : substeps;
    a.effects = effects; a.flags = flags;
    // staging of the CTA's prev_obs rows in shared memory: TMA bulk copy when every CTA's span is 16-byte aligned and sized,
    // per-thread LDGSTS otherwise; none when the span does not fit (e.g. 240 Hz control: 120-action buffers)
    {
        const size_t row_bytes = (size_t)a.obs_dim * 4, span = row_bytes * a.tpb;
        const bool aligned = aligned16(io->obs_prev) && aligned16(io->obs) && (span % 16 == 0) && ((row_bytes * ((size_t)a.N % a.tpb)) % 16 == 0);
        a.stage_rows = 0;
        if (io->obs && io->act_buffer_size > 0 && !state20 && span <= kStageLimit) a.stage_rows = (aligned && A == 4) ? 1 : 2;
    }
    if (state20) {
        const cudaError_t e2 = pid_act ? launch_step<true, true>(a, (cudaStream_t)stream) : launch_step<true, false>(a, (cudaStream_t)stream);
        return e2 == cudaSuccess ? 0 : cuda_fail(e2, "qs_step launch");
    }
    const cudaError_t e = pid_act ? launch_step<false, true>(a, (cudaStream_t)stream) : launch_step<false, false>(a, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : cuda_fail(e, "qs_step launch");
}

int qs_sizeof_rollout_io(void) { return (int)sizeof(QsRolloutIO); }

int qs_rollout_max_ticks(int act_type, int act_buffer_size, int drones_per_env) {
    const int A = act_width(act_type);
    if (A < 0 || act_type == QS_ACT_RAW_RPM || act_buffer_size <= 0 || drones_per_env <= 0 || drones_per_env > kMaxTPB) return 0;
    const size_t span = (size_t)block_size_for(drones_per_env) * (12 + act_buffer_size * A) * 4;
    if (span + 2 * (size_t)A * 4 > kStageLimit) return 0;
    return (int)((kStageLimit - span) / ((size_t)A * 4)) - 1;
}

int qs_rollout(const QsParams* p, const QsState* st, const QsRolloutIO* io, int act_type, int task,
               int n_envs, int drones_per_env, int substeps, unsigned effects, unsigned flags, void* stream) {
    if (!p || !io) return fail(QS_ERR_NULL, "qs_rollout: NULL params/io");
    if (int rc = check_state(st, (flags & QS_FLAG_AUTORESET_SAME_STEP) ? 1 : 0)) return rc;
    if (n_envs <= 0 || drones_per_env <= 0 || substeps <= 0 || io->T <= 0) return fail(QS_ERR_SIZE, "qs_rollout: sizes must be > 0");
    const int A = act_width(act_type);
    if (A < 0 || act_type == QS_ACT_RAW_RPM) return fail(QS_ERR_ENUM, "qs_rollout: bad act_type");
    if (task != QS_TASK_NONE && task != QS_TASK_HOVER) return fail(QS_ERR_ENUM, "qs_rollout: bad task");
    if (effects & ~7u) return fail(QS_ERR_ENUM, "qs_rollout: bad effects");
    if (flags & (QS_FLAG_AUTORESET_NEXT_STEP | QS_FLAG_SKIP_EPILOGUE | QS_FLAG_RPM_FROM_LAST | QS_FLAG_OBS_STATE20))
        return fail(QS_ERR_UNSUPPORTED, "qs_rollout: only SAME_STEP autoreset (or none) is supported");
    if (drones_per_env > kMaxTPB) return fail(QS_ERR_UNSUPPORTED, "qs_rollout: drones_per_env <= 128");
    if (!io->obs_init || !io->obs || !io->reward || !io->terminated || !io->truncated) return fail(QS_ERR_NULL, "qs_rollout: NULL buffer");
    if (io->act_buffer_size <= 0) return fail(QS_ERR_SIZE, "qs_rollout: act_buffer_size must be > 0");
    if (io->T > qs_rollout_max_ticks(act_type, io->act_buffer_size, drones_per_env)) return fail(QS_ERR_UNSUPPORTED, "qs_rollout: T exceeds qs_rollout_max_ticks (split the rollout)");
    if (task == QS_TASK_HOVER && (!st->target_pos || !aligned16(st->target_pos))) return fail(QS_ERR_NULL, "qs_rollout: target_pos NULL/misaligned");
    const bool pid_act = act_type == QS_ACT_PID || act_type == QS_ACT_VEL || act_type == QS_ACT_ONE_D_PID;
    if (pid_act && !st->pid) return fail(QS_ERR_NULL, "qs_rollout: PID action type needs QsState.pid");
    if ((effects & QS_EFFECT_DRAG) && !st->last_rpm) return fail(QS_ERR_NULL, "qs_rollout: DRAG needs QsState.last_rpm");
    if (A == 4 && ((io->actions && !aligned16(io->actions)) || (io->actions_out && !aligned16(io->actions_out)) || !aligned16(io->obs) || (io->obs_last && !aligned16(io->obs_last))))
        return fail(QS_ERR_ALIGN, "qs_rollout: [N][4]-wide buffers must be 16-byte aligned");
    RolloutArgs a;
    memset(&a, 0, sizeof(a));
    a.P = *p; a.st = *st; a.io = *io;
    a.act_type = act_type; a.task = task; a.n_envs = n_envs; a.D = drones_per_env; a.substeps = substeps;
    a.N = n_envs * drones_per_env; a.A = A; a.obs_dim = 12 + io->act_buffer_size * A;
    a.cap = cta_capacity(a.N, drones_per_env, true);
    a.tpb = block_size_for(drones_per_env, a.cap);
    a.effects = effects; a.flags = flags;
    {
        const size_t row_bytes = (size_t)a.obs_dim * 4, span = row_bytes * a.tpb;
        const bool aligned = aligned16(io->obs_init) && (span % 16 == 0) && ((row_bytes * ((size_t)a.N % a.tpb)) % 16 == 0);
        a.stage_mode = (aligned && A == 4) ? 1 : 2;
    }
    const int blocks = (int)((a.N + a.tpb - 1) / a.tpb);
    const int threads = ((a.tpb + 31) / 32) * 32;
    const size_t sm = smem_fixed(a.cap) + (size_t)a.tpb * a.obs_dim * 4 + (size_t)(io->T + 1) * A * 4 + 32;
    cudaStream_t s = (cudaStream_t)stream;
#define QS_RCASE(E)                                                                                                   \
    case E: {                                                                                                         \
        static bool set0 = false, set1 = false;                                                                       \
        if (pid_act) {                                                                                                \
            if (!set1) { cudaFuncSetAttribute(rollout_kernel<E, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kStepSmemFixed + kStageLimit + 32)); set1 = true; } \
            rollout_kernel<E, true><<<blocks, threads, sm, s>>>(a);                                                   \
        } else {                                                                                                      \
            if (!set0) { cudaFuncSetAttribute(rollout_kernel<E, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kStepSmemFixed + kStageLimit + 32)); set0 = true; } \
            rollout_kernel<E, false><<<blocks, threads, sm, s>>>(a);                                                  \
        }                                                                                                             \
    } break;
    switch (effects & 7u) { QS_RCASE(0) QS_RCASE(1) QS_RCASE(2) QS_RCASE(3) QS_RCASE(4) QS_RCASE(5) QS_RCASE(6) QS_RCASE(7) }
#undef QS_RCASE
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "qs_rollout launch");
}

int qs_sizeof_host_io(void) { return (int)sizeof(QsHostIO); }

int qs_step_call(const QsStepCall* c, void* stream) {
    if (!c) return fail(QS_ERR_NULL, "qs_step_call: NULL call");
    return qs_step(c->p, c->st, c->io, c->act_type, c->task, c->n_envs, c->drones_per_env, c->substeps, c->effects, c->flags, stream);
}

int qs_step_host(const QsParams* p, const QsState* st, const QsStepIO* io, const QsHostIO* h, int act_type, int task,
                 int n_envs, int drones_per_env, int substeps, unsigned effects, unsigned flags, void* stream) {
    if (!io || !h) return fail(QS_ERR_NULL, "qs_step_host: NULL io");
    if (!h->action_host || !h->obs_host || !h->reward_host || !h->terminated_host || !h->truncated_host || !h->action_dev)
        return fail(QS_ERR_NULL, "qs_step_host: NULL host buffer / action_dev");
    if (!io->obs || !io->reward || !io->terminated || !io->truncated) return fail(QS_ERR_NULL, "qs_step_host: NULL device buffer");
    const int A = act_width(act_type);
    if (A < 0) return fail(QS_ERR_ENUM, "qs_step_host: bad act_type");
    const bool state20 = flags & QS_FLAG_OBS_STATE20;
    const long long N = (long long)n_envs * drones_per_env;
    const int od = state20 ? 20 : 12 + io->act_buffer_size * A;
    const bool want_final = (flags & QS_FLAG_AUTORESET_SAME_STEP) && io->final_obs && h->final_obs_host;
    if (want_final && (!io->done || !h->done_host || !h->final_env_host || !h->n_final_host || !h->final_env_dev || !h->final_rows_dev))
        return fail(QS_ERR_NULL, "qs_step_host: final_obs transfer needs done / final_env / final_rows buffers");
    cudaStream_t s = (cudaStream_t)stream;
    static const bool trace = getenv("QS_TRACE") != nullptr;
    static int trace_n = 0;
    auto now = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; };
    const double t0 = trace ? now() : 0.0;
    double t1 = 0, t2 = 0, t3 = 0;
    cudaError_t e = cudaMemcpyAsync(h->action_dev, h->action_host, (size_t)N * A * 4, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return cuda_fail(e, "qs_step_host H2D action");
    QsStepIO dio = *io;
    dio.action = h->action_dev;
    if (int rc = qs_step(p, st, &dio, act_type, task, n_envs, drones_per_env, substeps, effects, flags, stream)) return rc;
    volatile int* marker = want_final ? h->n_final_host : nullptr;
    if (marker) *marker = -1;
    cudaMemcpyAsync(h->reward_host, io->reward, (size_t)n_envs * 4, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(h->terminated_host, io->terminated, (size_t)n_envs, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(h->truncated_host, io->truncated, (size_t)n_envs, cudaMemcpyDeviceToHost, s);
    if (io->done && h->done_host) cudaMemcpyAsync(h->done_host, io->done, (size_t)n_envs, cudaMemcpyDeviceToHost, s);
    if (want_final) {
        // The flags are 100 KB, the observations 19 MB.  A 4-byte copy queued behind the flag copies is their completion
        // marker (n_final_host, preset to -1: no reward has that bit pattern); the observation copy is queued right
        // after it, so it is in flight while the host waits for the marker, picks the finished aviaries and queues
        // their rows behind it.
        cudaMemcpyAsync(const_cast<int*>(marker), io->reward, 4, cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(h->obs_host, io->obs, (size_t)N * od * 4, cudaMemcpyDeviceToHost, s);
        if (trace) t1 = now();
        {
            const double t_spin = now();
            unsigned it = 0;
            while (*marker == -1) {
                if ((++it & 1023u) == 0 && now() - t_spin > 50000.0) {               // 50 ms: fall back to a full wait
                    e = cudaStreamSynchronize(s);
                    if (e != cudaSuccess) return cuda_fail(e, "qs_step_host sync(flags)");
                    break;
                }
            }
        }
        __atomic_thread_fence(__ATOMIC_ACQUIRE);                  // the flag arrays are read after the marker
        if (trace) t2 = now();
        int k = 0;
        for (int ev = 0; ev < n_envs; ++ev)
            if (h->done_host[ev]) h->final_env_host[k++] = ev;
        *h->n_final_host = k;
        if (k > 0) {
            const int row_floats = drones_per_env * od;
            cudaMemcpyAsync(h->final_env_dev, h->final_env_host, (size_t)k * 8, cudaMemcpyHostToDevice, s);
            gather_rows_kernel<<<k, 128, 0, s>>>(io->final_obs, h->final_env_dev, h->final_rows_dev, k, row_floats);
            cudaMemcpyAsync(h->final_obs_host, h->final_rows_dev, (size_t)k * row_floats * 4, cudaMemcpyDeviceToHost, s);
        }
    } else {
        if (h->n_final_host) *h->n_final_host = 0;
        cudaMemcpyAsync(h->obs_host, io->obs, (size_t)N * od * 4, cudaMemcpyDeviceToHost, s);
    }
    if (trace) t3 = now();
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return cuda_fail(e, "qs_step_host sync");
    if (trace && (++trace_n % 50) == 0)
        fprintf(stderr, "[qs_step_host] enqueue1 %.0f us, sync(flags) %.0f us, scan+enqueue2 %.0f us, final sync %.0f us, n_final %d\n",
                t1 - t0, t2 - t1, t3 - t2, now() - t3, h->n_final_host ? *h->n_final_host : -1);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "qs_step_host");
}

int qs_dyn_substeps(const QsParams* p, const QsState* st, const float* rpm, float* state20_out, const float* dw_fz,
                    int n_envs, int drones_per_env, int substeps, unsigned effects, unsigned flags, void* stream) {
    if (!p || !rpm) return fail(QS_ERR_NULL, "qs_dyn_substeps: NULL params/rpm");
    if (int rc = check_state(st, 0)) return rc;
    if (n_envs <= 0 || drones_per_env <= 0 || substeps <= 0) return fail(QS_ERR_SIZE, "qs_dyn_substeps: sizes must be > 0");
    if (!aligned16(rpm)) return fail(QS_ERR_ALIGN, "qs_dyn_substeps: rpm must be 16-byte aligned");
    if (effects & ~7u) return fail(QS_ERR_ENUM, "qs_dyn_substeps: bad effects");
    if ((effects & QS_EFFECT_DRAG) && !st->last_rpm) return fail(QS_ERR_NULL, "qs_dyn_substeps: DRAG needs QsState.last_rpm");
    if ((effects & QS_EFFECT_DW) && !dw_fz && drones_per_env > kMaxTPB)
        return fail(QS_ERR_UNSUPPORTED, "qs_dyn_substeps: in-CTA downwash needs drones_per_env <= 128 (else pass dw_fz, substeps = 1)");
    if ((effects & QS_EFFECT_DW) && dw_fz && substeps != 1) return fail(QS_ERR_UNSUPPORTED, "qs_dyn_substeps: external dw_fz requires substeps == 1");
    StepArgs a;
    memset(&a, 0, sizeof(a));
    a.P = *p; a.st = *st;
    a.io.action = rpm; a.io.obs = state20_out; a.io.dw_fz = dw_fz;
    a.act_type = QS_ACT_RAW_RPM; a.task = QS_TASK_NONE; a.n_envs = n_envs; a.D = drones_per_env; a.substeps = substeps;
    a.N = n_envs * drones_per_env; a.A = 4; a.obs_dim = 20;
    a.cap = cta_capacity(a.N, drones_per_env);
    a.tpb = block_size_for(drones_per_env, a.cap);
    a.counter_inc = substeps;
    a.effects = effects; a.flags = flags & QS_FLAG_RPY_F32;
    const cudaError_t e = launch_step<true, false>(a, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : cuda_fail(e, "qs_dyn_substeps launch");
}

int qs_pid_control(const QsParams* p, float* pid_state, double control_timestep,
                   const float* cur_pos, int pos_stride, const float* cur_quat, int quat_stride,
                   const float* cur_vel, int vel_stride,
                   const float* target_pos, const float* target_rpy, const float* target_vel, const float* target_rpy_rates,
                   int n, float* rpm_out, float* pos_e_out, float* yaw_e_out, void* stream) {
    if (!p || !pid_state || !cur_pos || !cur_quat || !cur_vel || !target_pos || !rpm_out) return fail(QS_ERR_NULL, "qs_pid_control: NULL argument");
    if (n <= 0 || pos_stride < 3 || quat_stride < 4 || vel_stride < 3) return fail(QS_ERR_SIZE, "qs_pid_control: bad n/stride");
    if (!(control_timestep > 0.0)) return fail(QS_ERR_SIZE, "qs_pid_control: control_timestep must be > 0");
    if (!aligned16(rpm_out)) return fail(QS_ERR_ALIGN, "qs_pid_control: rpm_out must be 16-byte aligned");
    PidArgs a;
    a.P = *p; a.pid = pid_state; a.dt = control_timestep;
    a.pos = cur_pos; a.quat = cur_quat; a.vel = cur_vel; a.tpos = target_pos; a.trpy = target_rpy; a.tvel = target_vel; a.trr = target_rpy_rates;
    a.pos_stride = pos_stride; a.quat_stride = quat_stride; a.vel_stride = vel_stride; a.n = n;
    a.rpm_out = rpm_out; a.pos_e_out = pos_e_out; a.yaw_e_out = yaw_e_out;
    pid_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "qs_pid_control launch");
}

static int launch_downwash(const QsParams* p, const float* rows, int n_envs, int D, float* fz_out, void* stream, const char* what) {
    const char* cull_env = getenv("QS_DW_CULL");              // QS_DW_CULL=0: evaluate every chunk (test / measurement switch)
    DwArgs a;
    a.prop_radius = (float)p->prop_radius; a.dw1 = (float)p->dw_coeff[0]; a.dw2 = (float)p->dw_coeff[1]; a.dw3 = (float)p->dw_coeff[2];
    a.rows = rows; a.fz = fz_out; a.D = D; a.tiles_per_env = (D + kDwDrones - 1) / kDwDrones;
    a.cull = (cull_env && cull_env[0] == '0') ? 0 : 1;
    downwash_kernel<<<n_envs * a.tiles_per_env, kDwDrones * kDwSlices, 0, (cudaStream_t)stream>>>(a);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, what);
}

int qs_downwash(const QsParams* p, const QsState* st, int n_envs, int drones_per_env, float* fz_out, void* stream) {
    if (!p || !st || !st->planes || !fz_out) return fail(QS_ERR_NULL, "qs_downwash: NULL argument");
    if (!aligned16(st->planes)) return fail(QS_ERR_ALIGN, "qs_downwash: planes must be 16-byte aligned");
    if (n_envs <= 0 || drones_per_env <= 0) return fail(QS_ERR_SIZE, "qs_downwash: sizes must be > 0");
    return launch_downwash(p, st->planes, n_envs, drones_per_env, fz_out, stream, "qs_downwash launch");
}

static int launch_downwash_boxed(const QsParams* p, const float* rows, int n_envs, int D, const float* src, int n_src, const float* boxes,
                                 const unsigned* ready, unsigned seq, int world, unsigned* err, float* fz_out, void* stream, const char* what) {
    const char* cull_env = getenv("QS_DW_CULL");
    DwbArgs a;
    a.prop_radius = (float)p->prop_radius; a.dw1 = (float)p->dw_coeff[0]; a.dw2 = (float)p->dw_coeff[1]; a.dw3 = (float)p->dw_coeff[2];
    a.rows = rows; a.fz = fz_out; a.src = src; a.boxes = boxes; a.n_src = n_src; a.D = D; a.groups_per_env = (D + 31) / 32;
    a.chunks = ((src ? n_src : D) + 31) / 32;
    a.cull = (cull_env && cull_env[0] == '0') ? 0 : 1; a.ready = ready; a.seq = seq; a.world = world; a.err = err;
    const long long blocks = (long long)n_envs * a.groups_per_env;
    if (blocks > 0x7fffffffLL) return fail(QS_ERR_SIZE, "downwash: too many row groups");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3(256); cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    // with flags: programmatic dependent of the publish kernel (which triggers at its first instruction): the launch
    // latency and the row loads overlap the push; the data dependency is carried by the flags, own rank's included
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = ready ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, downwash_boxed_kernel, a);
    return e == cudaSuccess ? 0 : cuda_fail(e, what);
}

static int launch_boxes(const float* pos, float* boxes, int n_envs, int D, void* stream, const char* what) {
    const int chunks = (D + 31) / 32;
    const long long total = (long long)n_envs * chunks;
    dw_boxes_kernel<<<(unsigned)((total + 7) / 8), 256, 0, (cudaStream_t)stream>>>(pos, boxes, D, chunks, total);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, what);
}

int qs_downwash_boxed(const QsParams* p, const QsState* st, int n_envs, int drones_per_env, float* boxes_ws, float* fz_out, void* stream) {
    if (!p || !st || !st->planes || !fz_out || !boxes_ws) return fail(QS_ERR_NULL, "qs_downwash_boxed: NULL argument");
    if (!aligned16(st->planes) || !aligned16(boxes_ws)) return fail(QS_ERR_ALIGN, "qs_downwash_boxed: planes / boxes_ws must be 16-byte aligned");
    if (n_envs <= 0 || drones_per_env <= 0) return fail(QS_ERR_SIZE, "qs_downwash_boxed: sizes must be > 0");
    if (int rc = launch_boxes(st->planes, boxes_ws, n_envs, drones_per_env, stream, "qs_downwash_boxed: boxes launch")) return rc;
    return launch_downwash_boxed(p, st->planes, n_envs, drones_per_env, nullptr, 0, boxes_ws, nullptr, 0u, 0, nullptr, fz_out, stream,
                                 "qs_downwash_boxed launch");
}

long long qs_dw_gathered_floats(int n_total) { return n_total > 0 ? 4LL * n_total + 8LL * ((n_total + 31) / 32) : 0; }

int qs_dw_boxes(float* gathered, int n_total, void* stream) {
    if (!gathered) return fail(QS_ERR_NULL, "qs_dw_boxes: NULL argument");
    if (!aligned16(gathered)) return fail(QS_ERR_ALIGN, "qs_dw_boxes: gathered must be 16-byte aligned");
    if (n_total <= 0) return fail(QS_ERR_SIZE, "qs_dw_boxes: n_total must be > 0");
    return launch_boxes(gathered, gathered + 4LL * n_total, 1, n_total, stream, "qs_dw_boxes launch");
}

int qs_downwash_rows(const QsParams* p, const float* rows_pos, int n_rows, const float* gathered, int n_total,
                     const unsigned* ready_flags, unsigned seq, int world, unsigned* err_flag, float* fz_out, void* stream) {
    if (!p || !rows_pos || !gathered || !fz_out) return fail(QS_ERR_NULL, "qs_downwash_rows: NULL argument");
    if (!aligned16(rows_pos) || !aligned16(gathered)) return fail(QS_ERR_ALIGN, "qs_downwash_rows: position arrays must be 16-byte aligned");
    if (n_rows <= 0 || n_total <= 0) return fail(QS_ERR_SIZE, "qs_downwash_rows: sizes must be > 0");
    if (ready_flags && (world <= 0 || world > QS_MAX_PEERS)) return fail(QS_ERR_SIZE, "qs_downwash_rows: world must be in [1, QS_MAX_PEERS]");
    return launch_downwash_boxed(p, rows_pos, 1, n_rows, gathered, n_total, gathered + 4LL * n_total, ready_flags, seq,
                                 ready_flags ? world : 0, err_flag, fz_out, stream, "qs_downwash_rows launch");
}

int qs_dw_publish(const float* pos, int n, int offset, float* const* gathered, int n_total, unsigned* const* flags, int world, int rank,
                  unsigned seq, unsigned* counter, void* stream) {
    if (!pos || !gathered || !flags || !counter) return fail(QS_ERR_NULL, "qs_dw_publish: NULL argument");
    if (world <= 0 || world > QS_MAX_PEERS || rank < 0 || rank >= world) return fail(QS_ERR_SIZE, "qs_dw_publish: bad world/rank");
    if (n <= 0 || offset < 0 || n_total < offset + n) return fail(QS_ERR_SIZE, "qs_dw_publish: bad n/offset/n_total");
    if (offset % 32 != 0 || (n % 32 != 0 && offset + n != n_total))
        return fail(QS_ERR_ALIGN, "qs_dw_publish: slices must start on a multiple of 32 drones (a chunk never straddles ranks)");
    if (!aligned16(pos)) return fail(QS_ERR_ALIGN, "qs_dw_publish: pos must be 16-byte aligned");
    PubArgs a;
    for (int r = 0; r < world; ++r) {
        if (!gathered[r] || !flags[r]) return fail(QS_ERR_NULL, "qs_dw_publish: NULL peer pointer");
        if (!aligned16(gathered[r])) return fail(QS_ERR_ALIGN, "qs_dw_publish: gathered arrays must be 16-byte aligned");
        a.dst[r] = gathered[r]; a.flags[r] = flags[r];
    }
    a.pos = pos; a.counter = counter; a.n = n; a.offset = offset; a.n_total = n_total; a.world = world; a.rank = rank; a.seq = seq;
    dw_publish_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "qs_dw_publish launch");
}

int qs_enable_peer_access(int peer_device) {
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); return 0; }
    return e == cudaSuccess ? 0 : cuda_fail(e, "qs_enable_peer_access");
}

// CUDA IPC of a (possibly sub-allocated) device buffer: handle of the enclosing allocation + byte offset.
int qs_ipc_export(const void* ptr, void* handle64, unsigned long long* offset) {
    if (!ptr || !handle64 || !offset) return fail(QS_ERR_NULL, "qs_ipc_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaError_t e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &qr);
    if (e != cudaSuccess || !f) return e != cudaSuccess ? cuda_fail(e, "qs_ipc_export: cuMemGetAddressRange lookup") : fail(QS_ERR_UNSUPPORTED, "qs_ipc_export: no cuMemGetAddressRange");
    unsigned long long base = 0;
    size_t size = 0;
    if (reinterpret_cast<range_fn>(f)(&base, &size, (unsigned long long)(uintptr_t)ptr) != 0) return fail(QS_ERR_UNSUPPORTED, "qs_ipc_export: cuMemGetAddressRange failed");
    e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), reinterpret_cast<void*>((uintptr_t)base));
    if (e != cudaSuccess) return cuda_fail(e, "qs_ipc_export: cudaIpcGetMemHandle");
    *offset = (unsigned long long)(uintptr_t)ptr - base;
    return 0;
}

int qs_ipc_import(const void* handle64, unsigned long long offset, void** ptr_out) {
    if (!handle64 || !ptr_out) return fail(QS_ERR_NULL, "qs_ipc_import: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* base = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return cuda_fail(e, "qs_ipc_import: cudaIpcOpenMemHandle");
    *ptr_out = static_cast<char*>(base) + offset;
    return 0;
}

int qs_adjacency(const QsState* st, int n_envs, int drones_per_env, double radius, unsigned char* out, void* stream) {
    if (!st || !st->planes || !out) return fail(QS_ERR_NULL, "qs_adjacency: NULL argument");
    if (!aligned16(st->planes)) return fail(QS_ERR_ALIGN, "qs_adjacency: planes must be 16-byte aligned");
    if (n_envs <= 0 || drones_per_env <= 0) return fail(QS_ERR_SIZE, "qs_adjacency: sizes must be > 0");
    if (drones_per_env % 16 == 0 && !aligned16(out)) return fail(QS_ERR_ALIGN, "qs_adjacency: out must be 16-byte aligned");
    AdjArgs a;
    a.planes = st->planes; a.out = out; a.radius = radius; a.D = drones_per_env;
    a.col_tiles = (drones_per_env + kAdjCols - 1) / kAdjCols; a.row_tiles = (drones_per_env + kAdjRows - 1) / kAdjRows;
    const long long blocks = (long long)n_envs * a.col_tiles * a.row_tiles;
    if (blocks > 0x7fffffffLL) return fail(QS_ERR_SIZE, "qs_adjacency: too many tiles");
    adjacency_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "qs_adjacency launch");
}

int qs_reset(const QsParams* p, const QsState* st, const unsigned char* mask, int n_envs, int drones_per_env,
             int reset_pid, float* obs, int obs_dim, int raw_state20, void* stream) {
    (void)p;
    if (int rc = check_state(st, 1)) return rc;
    if (n_envs <= 0 || drones_per_env <= 0) return fail(QS_ERR_SIZE, "qs_reset: sizes must be > 0");
    if (obs && !raw_state20 && obs_dim < 12) return fail(QS_ERR_SIZE, "qs_reset: obs_dim < 12");
    ResetArgs a;
    a.st = *st; a.mask = mask; a.D = drones_per_env; a.reset_pid = reset_pid; a.obs_dim = obs_dim; a.raw20 = raw_state20;
    a.N = (long long)n_envs * drones_per_env; a.obs = obs;
    reset_kernel<<<(int)((a.N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "qs_reset launch");
}

}  // extern "C"
