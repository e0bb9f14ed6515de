// rollout.cu -- qs_rollout: T fused control ticks per launch.
#include "qs_common.cuh"

using namespace qsi;

namespace {

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
    QsPolicy pol;        // copy of *io.policy (device pointers inside) when POLICY
    int act_type, task, n_envs, D, substeps, N, A, obs_dim, tpb;
    unsigned effects, flags;
    int stage_mode, cap;
};

// ---- on-device policy: SB3-MlpPolicy-shaped MLP evaluated by ONE WARP per network (warp 0 actor, warp 1 critic) ---------------
constexpr int kHid = 64, kHidStride = 68;      // hidden rows padded to 68 floats: conflict-free 128-bit reads across 8 rows

__device__ __forceinline__ float4 ld4_any(const float* p, bool vec) {
    if (vec) return *reinterpret_cast<const float4*>(p);
    return make_float4(p[0], p[1], p[2], p[3]);
}

// y[av][0..63] = act(b + W^T x[av]) for av < n_av (n_av <= 32).  x rows: K floats at stride x_stride in shared memory; W [K][64]
// row-major in global memory (read through L1: every lane group of a warp and every CTA of the SM reads the same 8 KB..37 KB).
// Lane = (aviary group ag = lane / 8, unit group ug = lane % 8): 8 aviaries (ag + 4 j) x 8 units (8 ug + u) = 64 accumulators;
// per 4 k: 8 x-loads + 8 weight loads for 256 FFMA.  May run in place (y_s == x_s, same stride): all reads precede all writes.
template <bool TANH>
__device__ __forceinline__ void mlp_layer64(const float* x_s, int x_stride, int K, bool x_vec, const float* __restrict__ W,
                                            const float* __restrict__ b, float* y_s, int n_av, int lane) {
    const int ag = lane >> 3, ug = lane & 7;
    float acc[8][8];
    {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + 8 * ug)), b1 = __ldg(reinterpret_cast<const float4*>(b + 8 * ug + 4));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            acc[j][0] = b0.x; acc[j][1] = b0.y; acc[j][2] = b0.z; acc[j][3] = b0.w;
            acc[j][4] = b1.x; acc[j][5] = b1.y; acc[j][6] = b1.z; acc[j][7] = b1.w;
        }
    }
    const int K4 = K & ~3;
    for (int k = 0; k < K4; k += 4) {
        float4 xv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int av = ag + 4 * j;
            xv[j] = av < n_av ? ld4_any(x_s + (size_t)av * x_stride + k, x_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(W + (size_t)(k + kk) * kHid + 8 * ug));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(W + (size_t)(k + kk) * kHid + 8 * ug + 4));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float x = kk == 0 ? xv[j].x : (kk == 1 ? xv[j].y : (kk == 2 ? xv[j].z : xv[j].w));
                acc[j][0] = fmaf(x, w0.x, acc[j][0]); acc[j][1] = fmaf(x, w0.y, acc[j][1]);
                acc[j][2] = fmaf(x, w0.z, acc[j][2]); acc[j][3] = fmaf(x, w0.w, acc[j][3]);
                acc[j][4] = fmaf(x, w1.x, acc[j][4]); acc[j][5] = fmaf(x, w1.y, acc[j][5]);
                acc[j][6] = fmaf(x, w1.z, acc[j][6]); acc[j][7] = fmaf(x, w1.w, acc[j][7]);
            }
        }
    }
    for (int k = K4; k < K; ++k) {                        // K not a multiple of 4 (A = 1 observations)
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(W + (size_t)k * kHid + 8 * ug));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(W + (size_t)k * kHid + 8 * ug + 4));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int av = ag + 4 * j;
            const float x = av < n_av ? x_s[(size_t)av * x_stride + k] : 0.f;
            acc[j][0] = fmaf(x, w0.x, acc[j][0]); acc[j][1] = fmaf(x, w0.y, acc[j][1]);
            acc[j][2] = fmaf(x, w0.z, acc[j][2]); acc[j][3] = fmaf(x, w0.w, acc[j][3]);
            acc[j][4] = fmaf(x, w1.x, acc[j][4]); acc[j][5] = fmaf(x, w1.y, acc[j][5]);
            acc[j][6] = fmaf(x, w1.z, acc[j][6]); acc[j][7] = fmaf(x, w1.w, acc[j][7]);
        }
    }
    __syncwarp();                                         // in-place layers: every read of x precedes the first write of y
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int av = ag + 4 * j;
        if (av < n_av) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = TANH ? tanhf(acc[j][u]) : acc[j][u];
            float4* y = reinterpret_cast<float4*>(y_s + (size_t)av * kHidStride + 8 * ug);
            y[0] = make_float4(v[0], v[1], v[2], v[3]); y[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
    __syncwarp();
}

// out[av][j] = b[j] + sum_k h[av][k] W[k][j], j < n_out: lane = aviary, the weights are warp-uniform loads
__device__ __forceinline__ void mlp_head(const float* h_s, const float* __restrict__ W, const float* __restrict__ b, int n_out,
                                         float* out_s, int n_av, int lane) {
    if (lane >= n_av) return;
    const float* h = h_s + (size_t)lane * kHidStride;
    for (int j0 = 0; j0 < n_out; j0 += 4) {
        const int nj = n_out - j0 < 4 ? n_out - j0 : 4;
        float a0 = __ldg(b + j0), a1 = nj > 1 ? __ldg(b + j0 + 1) : 0.f, a2 = nj > 2 ? __ldg(b + j0 + 2) : 0.f, a3 = nj > 3 ? __ldg(b + j0 + 3) : 0.f;
        for (int k = 0; k < kHid; k += 4) {
            const float4 hv = *reinterpret_cast<const float4*>(h + k);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float x = kk == 0 ? hv.x : (kk == 1 ? hv.y : (kk == 2 ? hv.z : hv.w));
                const float* w = W + (size_t)(k + kk) * n_out + j0;
                a0 = fmaf(x, __ldg(w), a0);
                if (nj > 1) a1 = fmaf(x, __ldg(w + 1), a1);
                if (nj > 2) a2 = fmaf(x, __ldg(w + 2), a2);
                if (nj > 3) a3 = fmaf(x, __ldg(w + 3), a3);
            }
        }
        float* o = out_s + (size_t)lane * n_out + j0;
        o[0] = a0; if (nj > 1) o[1] = a1; if (nj > 2) o[2] = a2; if (nj > 3) o[3] = a3;
    }
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float u32_to_pm1(unsigned u) { return (float)(u >> 8) * (1.0f / 8388608.0f) - 1.0f; }   // [-1, 1)

// what store_drone + load_drone do to the state between two ticks: the quaternion is renormalised, nothing is rounded
// (the planes are float64), so T fused ticks equal T calls of qs_step bit for bit
__device__ __forceinline__ void round_to_planes(qs::Drone& d) {
    const double inv = rsqrt(qs::quat_norm2(d.qx, d.qy, d.qz, d.qw));
    d.qx = __dmul_rn(d.qx, inv); d.qy = __dmul_rn(d.qy, inv); d.qz = __dmul_rn(d.qz, inv); d.qw = __dmul_rn(d.qw, inv);
}

template <int EFF, bool PIDACT, bool POLICY>
__global__ void __launch_bounds__(POLICY ? 64 : kMaxTPB, POLICY ? 3 : 4) rollout_kernel(const __grid_constant__ RolloutArgs a) {
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
    // POLICY: [2 nets][32 aviaries][68] hidden activations, [32][out_dim] means, [32] values, [tpb] log-prob terms, after the window
    float* pol_s = stage_s + ((((size_t)tpb * od + (size_t)(T + 1) * A) + 3) & ~(size_t)3);
    float* mean_s = pol_s + 2 * 32 * kHidStride;                                        // [<= 128 aviaries][out_dim]
    float* val_s = mean_s + (size_t)kMaxTPB * a.pol.out_dim;
    float* lp_s = val_s + kMaxTPB;

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
        if ((EFF & QS_EFFECT_DRAG) && a.st.last_rpm) load_rpm(a.st.last_rpm, i, rpm_prev);
        if (PIDACT) load_pid(a.st.pid, N, i, pst);
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
        float raw_act[4] = {0.f, 0.f, 0.f, 0.f};
        if (POLICY) {
            // the aviaries of this CTA: rows [le D, le D + D) of the window = one flattened observation of in_dim floats each
            const int n_av = rows / D, warp = t >> 5, lane = t & 31;
            const bool x_vec = (A == 4) && ((a.pol.in_dim & 3) == 0);
            for (int a0 = 0; a0 < n_av; a0 += 32) {           // 32 aviaries per pass (one pass for D >= 2)
                const int na = n_av - a0 < 32 ? n_av - a0 : 32;
                const float* x0 = base + (size_t)a0 * a.pol.in_dim;
                if (warp == 0) {
                    float* h = pol_s;
                    mlp_layer64<true>(x0, a.pol.in_dim, a.pol.in_dim, x_vec, a.pol.w1, a.pol.b1, h, na, lane);
                    mlp_layer64<true>(h, kHidStride, kHid, true, a.pol.w2, a.pol.b2, h, na, lane);
                    mlp_head(h, a.pol.w3, a.pol.b3, a.pol.out_dim, mean_s + (size_t)a0 * a.pol.out_dim, na, lane);
                } else if (warp == 1 && a.pol.vw1) {
                    float* h = pol_s + 32 * kHidStride;
                    mlp_layer64<true>(x0, a.pol.in_dim, a.pol.in_dim, x_vec, a.pol.vw1, a.pol.vb1, h, na, lane);
                    mlp_layer64<true>(h, kHidStride, kHid, true, a.pol.vw2, a.pol.vb2, h, na, lane);
                    mlp_head(h, a.pol.vw3, a.pol.vb3, 1, val_s + a0, na, lane);
                }
            }
            __syncthreads();
            float lp = 0.f;
            if (live) {
                const int od_out = a.pol.out_dim;
                for (int j = 0; j < A; ++j) {
                    const int idx = dslot * A + j;
                    const float ls = __ldg(a.pol.log_std + idx);
                    const float eps = a.pol.noise ? __ldg(a.pol.noise + ((long long)k * E + e) * od_out + idx) : 0.f;
                    const float r = fmaf(expf(ls), eps, mean_s[le * od_out + idx]);
                    raw_act[j] = r;
                    act[j] = fminf(fmaxf(r, -1.f), 1.f);                             // the env clips to its action space
                    lp += -0.5f * eps * eps - ls - 0.91893853320467274f;            // log N(r; mean, std)
                }
                lp_s[t] = lp;
            }
            __syncthreads();
            if (live && dslot == 0) {
                float s_lp = 0.f;
                for (int q = 0; q < D; ++q) s_lp += lp_s[t + q];
                const long long oe = (long long)k * E + e;
                if (a.pol.logprob) a.pol.logprob[oe] = s_lp;
                if (a.pol.values && a.pol.vw1) a.pol.values[oe] = val_s[le];
            }
        }
        if (live) {
            if (POLICY) {
                // (act / raw_act set above)
            } else if (a.io.actions) {
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
                const float* av = POLICY ? raw_act : act;                        // PPO stores the unclipped sample
                if (A == 4) *reinterpret_cast<float4*>(ao) = make_float4(av[0], av[1], av[2], av[3]);
                else if (A == 3) { ao[0] = av[0]; ao[1] = av[1]; ao[2] = av[2]; }
                else ao[0] = av[0];
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
                const D4 tp = ld256_nc(a.st.target_pos, tbl);
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
            rpm_prev[0] = rpm[0]; rpm_prev[1] = rpm[1]; rpm_prev[2] = rpm[2]; rpm_prev[3] = rpm[3];
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
        store_drone(a.st, N, i, d);
        if (a.st.last_rpm) st256(a.st.last_rpm, i, rpm[0], rpm[1], rpm[2], rpm[3]);
        if (PIDACT) store_pid(a.st.pid, N, i, pst);
        if (dslot == 0) a.st.step_counter[e] = sc;
    }
}

}  // namespace

extern "C" {

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
    if (task == QS_TASK_HOVER && (!st->target_pos || !aligned32(st->target_pos))) return fail(QS_ERR_NULL, "qs_rollout: target_pos NULL/misaligned");
    const bool pid_act = act_type == QS_ACT_PID || act_type == QS_ACT_VEL || act_type == QS_ACT_ONE_D_PID;
    if (pid_act && !st->pid) return fail(QS_ERR_NULL, "qs_rollout: PID action type needs QsState.pid");
    if ((effects & QS_EFFECT_DRAG) && !st->last_rpm) return fail(QS_ERR_NULL, "qs_rollout: DRAG needs QsState.last_rpm");
    if (A == 4 && ((io->actions && !aligned16(io->actions)) || (io->actions_out && !aligned16(io->actions_out)) || !aligned16(io->obs) || (io->obs_last && !aligned16(io->obs_last))))
        return fail(QS_ERR_ALIGN, "qs_rollout: [N][4]-wide buffers must be 16-byte aligned");
    RolloutArgs a;
    memset(&a, 0, sizeof(a));
    a.P = *p; a.st = *st; a.io = *io;
    a.act_type = act_type; a.task = task; a.n_envs = n_envs; a.D = drones_per_env; a.substeps = substeps;
    if ((long long)n_envs * drones_per_env > 0x7fffffffLL) return fail(QS_ERR_SIZE, "qs_rollout: n_envs * drones_per_env exceeds 2^31-1");
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
    int threads = ((a.tpb + 31) / 32) * 32;
    size_t sm = smem_fixed(a.cap) + (size_t)a.tpb * a.obs_dim * 4 + (size_t)(io->T + 1) * A * 4 + 32;
    cudaStream_t s = (cudaStream_t)stream;
    if (io->policy) {
        const QsPolicy& q = *io->policy;
        if (pid_act || (effects & 7u) || a.cap > 64) return fail(QS_ERR_UNSUPPORTED, "qs_rollout: the on-device policy supports RPM / ONE_D_RPM actions, no DYN+ effects, drones_per_env <= 64");
        if (io->actions) return fail(QS_ERR_UNSUPPORTED, "qs_rollout: pass either actions or a policy");
        if (!q.w1 || !q.b1 || !q.w2 || !q.b2 || !q.w3 || !q.b3 || !q.log_std) return fail(QS_ERR_NULL, "qs_rollout: policy weights are NULL");
        if (q.in_dim != drones_per_env * a.obs_dim || q.out_dim != drones_per_env * A) return fail(QS_ERR_SIZE, "qs_rollout: policy in_dim/out_dim must be D*obs_dim / D*A");
        if (q.vw1 && (!q.vb1 || !q.vw2 || !q.vb2 || !q.vw3 || !q.vb3)) return fail(QS_ERR_NULL, "qs_rollout: incomplete critic");
        if (q.values && !q.vw1) return fail(QS_ERR_NULL, "qs_rollout: values requested without a critic");
        for (const float* w : {q.w1, q.b1, q.w2, q.b2, q.vw1, q.vb1, q.vw2, q.vb2})
            if (w && !aligned16(w)) return fail(QS_ERR_ALIGN, "qs_rollout: policy matrices must be 16-byte aligned");
        a.pol = q;
        threads = 64;                                            // warp 0: actor, warp 1: critic
        sm += 16 + (size_t)(2 * 32 * kHidStride + kMaxTPB * q.out_dim + 2 * kMaxTPB) * 4;
        if (sm > 200 * 1024) return fail(QS_ERR_UNSUPPORTED, "qs_rollout: policy + window exceed shared memory");
        if (sm > 48 * 1024) cudaFuncSetAttribute(rollout_kernel<0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        rollout_kernel<0, false, true><<<blocks, threads, sm, s>>>(a);
        const cudaError_t e = cudaGetLastError();
        return e == cudaSuccess ? 0 : cuda_fail(e, "qs_rollout (policy) launch");
    }
#define QS_RCASE(E)                                                                                                   \
    case E: {                                                                                                         \
        if (pid_act) {                                                                                                \
            if (sm > 48 * 1024) cudaFuncSetAttribute(rollout_kernel<E, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kStepSmemFixed + kStageLimit + 32)); \
            rollout_kernel<E, true, false><<<blocks, threads, sm, s>>>(a);                                            \
        } else {                                                                                                      \
            if (sm > 48 * 1024) cudaFuncSetAttribute(rollout_kernel<E, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kStepSmemFixed + kStageLimit + 32)); \
            rollout_kernel<E, false, false><<<blocks, threads, sm, s>>>(a);                                           \
        }                                                                                                             \
    } break;
    switch (effects & 7u) { QS_RCASE(0) QS_RCASE(1) QS_RCASE(2) QS_RCASE(3) QS_RCASE(4) QS_RCASE(5) QS_RCASE(6) QS_RCASE(7) }
#undef QS_RCASE
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "qs_rollout launch");
}

}  // extern "C"
