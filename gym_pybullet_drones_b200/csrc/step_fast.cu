// step_fast.cu -- the fused control tick for the common RL configurations (BaseAviary.step, envs/BaseAviary.py:259-383,
// with BaseRLAviary's RPM / ONE_D_RPM actions, KIN observations and the Hover/MultiHover task): no DYN+ effects, no
// embedded PID, aviaries of 1, 2, 4, ... 32 drones, autoreset SAME_STEP or none.  Everything else takes step_general.cu.
//
// Same arithmetic as the general kernel (the per-drone functions of quad_core.cuh), different skeleton:
//   * a WARP is the unit of work: 32 consecutive drones = one contiguous span of observation rows with its own mbarrier
//     and shared-memory window; no __syncthreads anywhere, the aviary reduction is a warp shuffle (aviaries never straddle
//     warps), so warps of a CTA run completely out of phase
//   * A, the task, the autoreset mode and the rpy precision are template parameters: the instruction stream of an
//     instantiation contains no mode switches (round 1's kernel: 268 IMAD, 101 BRA per warp)
//   * state in/out as 3 x 32-byte + 1 x 8-byte accesses per thread (float64 planes, no conversions)
//   * A = 4: the old span is TMA-loaded, patched in place (head -> slots [A, A+12), new action -> the A slots after the
//     row) and TMA-stored shifted by one action (16 bytes); A = 1: the span is TMA-loaded, funnel-shifted by one float with
//     128-bit shared-memory accesses into a second 16-byte-aligned window, patched there and TMA-stored
//   * the observation of a freshly reset drone (SAME_STEP autoreset) comes from a precomputed table (qs_reset_heads),
//     not from two atan2f and an asinf in the epilogue
#include "qs_common.cuh"

#ifdef QS_TIMELINE
// debug build only (tools/timeline.py): per-warp phase timestamps (%globaltimer, ns) of the last launch
__device__ unsigned long long g_timeline[4][8192 * 16];
static int g_dbg_slot = 0;
#define QS_STAMP(k) do { if (lane == 0 && wg < 8192) g_timeline[a.dbg_slot & 3][wg * 16 + (k)] = globaltimer_ns(); } while (0)
#else
#define QS_STAMP(k) do { } while (0)
#endif

namespace qsi {
namespace {

template <int A>
struct FastSmem {
    // per-warp shared-memory window, in floats
    static __host__ __device__ constexpr int span(int od) { return 32 * od; }
    // A = 4: [span + 4 tail floats (rounded to 16 B)] ; A = 1: [X: span + 4 (the shift reads one float past the span)] [Y: span]
    static __host__ __device__ constexpr int x_floats(int od) { return (span(od) + 4 + 3) / 4 * 4; }
    static __host__ __device__ constexpr int y_floats(int od) { return A == 4 ? 0 : (span(od) + 3) / 4 * 4; }
    static __host__ __device__ constexpr int fin_floats(bool fin) { return fin ? 32 * 12 : 0; }
    static __host__ __device__ constexpr int total_bytes(int od, bool fin) {
        return ((x_floats(od) + y_floats(od) + fin_floats(fin)) * 4 + 16 + 127) / 128 * 128;      // + mbarrier, 128-byte multiple
    }
};

// A: action width (4 = RPM, 1 = ONE_D_RPM).  TASK: Hover/MultiHover reward + flags (else the CtrlAviary-style dummy task).
// RESET: SAME_STEP autoreset.  RPYF: float32 atan2f/asinf for the reported rpy.  WARPS: warps per CTA (independent).
template <int A, bool TASK, bool RESET, bool RPYF, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) step_fast_kernel(const __grid_constant__ StepArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const QsParams& P = a.P;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int od = a.obs_dim;
    const long long N = a.N;
    const long long wg = (long long)a.first_warp + (long long)blockIdx.x * WARPS + warp;      // global warp index (a launch may cover a chunk)
    const long long w0 = wg * 32;                                           // first drone of this warp
    if (w0 >= N) return;                                                    // (whole warp: no barrier is shared between warps)
    if (WARPS > 1 && a.n_warps > 0 && wg >= (long long)a.first_warp + a.n_warps) return;     // past the chunk (a later launch owns these drones)
    const long long i = w0 + lane;
    const bool live = i < N;
    const int rows = (int)((N - w0) < 32 ? (N - w0) : 32);
    const bool want_fin = RESET && a.io.final_obs != nullptr;

    float* xs = reinterpret_cast<float*>(smem_raw + (size_t)warp * FastSmem<A>::total_bytes(od, want_fin));
    float* ys = xs + FastSmem<A>::x_floats(od);
    float* fin_s = ys + FastSmem<A>::y_floats(od);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(fin_s + FastSmem<A>::fin_floats(want_fin));

    const int D = a.D, dmask = D - 1;                                       // D is a power of two <= 32
    const long long il = live ? i : w0;                                     // dead lanes of a ragged last warp shadow the first drone
    const long long e = il >> a.log2D;
    const int dslot = (int)il & dmask;
    const long long tbl = a.st.tables_per_env ? il : dslot;

    const float* span_src = a.io.obs_prev + w0 * od;
    float* span_dst = a.io.obs + w0 * od;
    const unsigned span_bytes = (unsigned)(rows * od * 4);
    QS_STAMP(0);
    const bool hint_writer = a.io.pdl_hint != nullptr && wg == 0 && lane == 0;
    unsigned long long t_wait0 = 0;
    if (hint_writer) t_wait0 = globaltimer_ns();
    if (lane == 0) {
        mbar_init(bar, 1);
        if (a.prefetch) {
            // Programmatic dependent launch: this CTA may be resident while the previous kernel of the stream is still in its
            // compute / store phases with the memory system idle.  Pulling this warp's inputs into L2 now is always safe (L2 is
            // the point of coherence: lines the previous kernel still writes are simply updated) and turns the DRAM round trips
            // after the dependency wait into L2 hits.
            auto pf = [](const void* p, unsigned bytes) {
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
            };
            const unsigned nb = (unsigned)rows * 32u;
            pf(a.st.planes + 4 * w0, nb); pf(a.st.planes + 4 * (N + w0), nb); pf(a.st.planes + 4 * (2 * N + w0), nb);
            if (((rows * 8) & 15) == 0) pf(a.st.planes + 12 * N + w0, (unsigned)rows * 8u);
            if (((rows * A * 4) & 15) == 0) pf(a.io.action + w0 * A, (unsigned)(rows * A * 4));
            // the history too: back-to-back launches then find it in L2 and write it back during the FP64 loop (-0.5 us per step);
            // a launch on an idle GPU has no window and would pay ~2 us because its state loads queue behind it (measured), so
            // the previous launch on these buffers says whether it saw a window (pdl_hint)
            if (a.prefetch > 1 && (a.io.pdl_hint == nullptr || *reinterpret_cast<const volatile unsigned*>(a.io.pdl_hint) != 0u)) pf(span_src, span_bytes);
        }
    }
    __syncwarp();
    // read-only tables (never written by a kernel): safe ahead of the dependency wait
    double tpx = 0.0, tpy = 0.0, tpz = 0.0;
    if (TASK) { const D4 tp = ld256_nc(a.st.target_pos, tbl); tpx = tp.x; tpy = tp.y; tpz = tp.z; }

    // nothing written by the previous kernel in the stream is read above this line (programmatic dependent launch)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");        // let the next grid's CTAs take the free slots now
    if (hint_writer) *a.io.pdl_hint = (globaltimer_ns() - t_wait0 > 1500ull) ? 1u : 0u;      // was I resident > 1.5 us before my dependency resolved?
    QS_STAMP(1);

    // ---- loads: state (3 x 32 B + 8 B), action, step counter; then the bulk copy of the old observation span ------------
    qs::Drone d;
    float act[4] = {0.f, 0.f, 0.f, 0.f};
    int sc = 0;
    load_drone(a.st.planes, N, il, d);
    if (A == 4) {
        const float4 v = ldg4(a.io.action, il);
        act[0] = v.x; act[1] = v.y; act[2] = v.z; act[3] = v.w;
    } else {
        act[0] = __ldg(a.io.action + il);
    }
    sc = a.st.step_counter[e];
    // The bulk copy of the old span (9 KB per warp) is issued only once the step counter -- and with it the batch of small
    // state loads issued just before it -- has ARRIVED: warps issue in order, so the comparison below stalls until then, and
    // the memory system serves every warp's 120 bytes of state ahead of the 19 MB of history the physics does not need yet.
    // (The comparison is always true for a valid counter; the compiler cannot know.)
    // Experiment (QS_ROW_LOADS=1, off): every lane fetches only the history of its own row, leaving the 48-byte head and the
    // dropped oldest action (64 of 288 bytes per row) in HBM.  32 small bulk copies per warp measured SLOWER than one copy of
    // the whole span (13.2 vs 12.3 us per step): the default moves the span.
    const bool by_rows = A == 4 && a.row_loads;
    auto load_span = [&]() {
        if (by_rows) {
            const unsigned hb = (unsigned)(od - 16) * 4u;
            if (lane == 0) mbar_expect_tx(bar, hb * (unsigned)rows);
            __syncwarp();
            if (live) bulk_g2s(xs + (size_t)lane * od + 16, span_src + (size_t)lane * od + 16, hb, bar);
        } else if (lane == 0) {
            tma_bulk_g2s(xs, span_src, span_bytes, bar);
        }
    };
    bool issued = false;
    if (a.flags_late_tma == 0 || __shfl_sync(0xffffffffu, sc, 0) != (int)0x80000000) {
        load_span();
        issued = true;
    }
    QS_STAMP(2);

    // ---- action decode (BaseRLAviary.py:192,225) + S substeps ---------------------------------------------------------------
    double rpm[4];
    {
        qs::PidState none = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        qs::decode_action<false>(P, A == 4 ? QS_ACT_RPM : QS_ACT_ONE_D_RPM, act, d, 0.0, none, rpm);
    }
    // A = 4: the history part of the new rows does not depend on the physics: as soon as the old span has landed (polled between
    // substeps) the copy engine writes it back shifted by one action; the heads and the new actions follow at the end as
    // ordinary stores.  So the 19 MB of history stores overlap the FP64 loop instead of following it.
    const float4* shifted = reinterpret_cast<const float4*>(xs) + 1;
    bool stored = false;
    auto store_span = [&]() {
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(span_dst), "r"(smem_u32(shifted)), "r"(span_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        stored = true;
    };
    auto poll = [&](int s) {
        if (A == 4 && a.early_store && issued && !stored && (s & 1)) {
            int ok = 0;
            if (lane == 0) ok = mbar_test(bar, 0) ? 1 : 0;
            if (__shfl_sync(0xffffffffu, ok, 0)) store_span();
        }
    };
    double R_last[9];
    qs::dyn_tick<0>(P, d, rpm, rpm, 0.0, a.substeps, R_last, poll);
    QS_STAMP(3);
    qs::Derived o;
    qs::derive<RPYF>(d, R_last, o);
    QS_STAMP(4);

    // ---- task terms, reduced over the D drones of the aviary in index order (MultiHoverAviary.py:75-130) ----------------------
    bool env_done = false;
    float g_rew = -1.0f;
    bool g_term = false, g_trunc = false;
    if (TASK) {
        const qs::TaskTerms tt = qs::hover_terms(P, d, o, tpx, tpy, tpz);
        double rew = 0.0, dist = 0.0;
        const int base = lane & ~dmask;
        for (int k = 0; k < D; ++k) {
            rew += __shfl_sync(0xffffffffu, tt.reward, base + k);
            dist += __shfl_sync(0xffffffffu, tt.dist, base + k);
        }
        const unsigned oobs = __ballot_sync(0xffffffffu, tt.out_of_bounds && live);
        const unsigned gmask = (D == 32 ? 0xffffffffu : ((1u << D) - 1u)) << base;
        const bool term = dist < P.term_dist;                                     // HoverAviary.py:91
        const bool trunc = (oobs & gmask) != 0u || sc >= a.sc_limit;              // HoverAviary.py:113 (sc/PYB_FREQ > EPISODE_LEN_SEC)
        env_done = term || trunc;
        g_rew = (float)rew; g_term = term; g_trunc = trunc;
        if (live && dslot == 0) {
            a.io.reward[e] = (float)rew;
            a.io.terminated[e] = term ? 1 : 0;
            a.io.truncated[e] = trunc ? 1 : 0;
            if (a.io.done) a.io.done[e] = env_done ? 1 : 0;
        }
    } else if (live && dslot == 0) {
        a.io.reward[e] = -1.0f; a.io.terminated[e] = 0; a.io.truncated[e] = 0;     // CtrlAviary-style dummy task
        if (a.io.done) a.io.done[e] = 0;
    }
    QS_STAMP(5);

    // ---- observation head, autoreset, state store ----------------------------------------------------------------------------
    float h[12];
    h[0] = (float)d.px; h[1] = (float)d.py; h[2] = (float)d.pz;                    // BaseRLAviary.py:310-315
    h[3] = (float)o.roll; h[4] = (float)o.pitch; h[5] = (float)o.yaw;
    h[6] = (float)d.vx; h[7] = (float)d.vy; h[8] = (float)d.vz;
    h[9] = (float)o.ax; h[10] = (float)o.ay; h[11] = (float)o.az;
    const bool reset_me = RESET && env_done;
    if (reset_me) {
        if (want_fin) {                                                            // terminal head, for final_obs
            float4* f4 = reinterpret_cast<float4*>(fin_s + 12 * lane);
            f4[0] = make_float4(h[0], h[1], h[2], h[3]); f4[1] = make_float4(h[4], h[5], h[6], h[7]); f4[2] = make_float4(h[8], h[9], h[10], h[11]);
        }
        init_drone(a.st, tbl, d);                                                  // BaseAviary.py:451-505
        const float4* rh = reinterpret_cast<const float4*>(a.st.reset_head) + 3 * tbl;
        const float4 r0 = __ldg(rh), r1 = __ldg(rh + 1), r2 = __ldg(rh + 2);
        h[0] = r0.x; h[1] = r0.y; h[2] = r0.z; h[3] = r0.w; h[4] = r1.x; h[5] = r1.y; h[6] = r1.z; h[7] = r1.w;
        h[8] = r2.x; h[9] = r2.y; h[10] = r2.z; h[11] = r2.w;
        rpm[0] = rpm[1] = rpm[2] = rpm[3] = 0.0;                                   // last_clipped_action = 0
        sc = -a.counter_inc;
    }
    if (live) {
        store_drone(a.st, N, i, d);
        if (a.st.last_rpm) st256(a.st.last_rpm, i, rpm[0], rpm[1], rpm[2], rpm[3]);
        if (dslot == 0) a.st.step_counter[e] = sc + a.counter_inc;                 // BaseAviary.py:382
    }
    QS_STAMP(6);

    // ---- observation rows --------------------------------------------------------------------------------------------------------
    if (!issued) load_span();
    const unsigned fin_rows = want_fin ? __ballot_sync(0xffffffffu, reset_me && live) : 0u;
    // Fused observation gather: the finished rows go a second time, straight from shared memory, to the learner's tensor
    // (peer memory over NVLink), the per-aviary outputs with them; the last warp of the grid raises the learner's flag.
    auto gather = [&](const void* rows_smem) {
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(a.io.obs_gather + w0 * od), "r"(smem_u32(rows_smem)), "r"(span_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (live && dslot == 0) {
            if (a.io.reward_gather) a.io.reward_gather[e] = g_rew;
            if (a.io.terminated_gather) a.io.terminated_gather[e] = g_term ? 1 : 0;
            if (a.io.truncated_gather) a.io.truncated_gather[e] = g_trunc ? 1 : 0;
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // rows written (not only read) before the flag
        __syncwarp();
        if (a.io.gather_flag) {
            __threadfence_system();
            if (lane == 0) {
                const unsigned nwarps = (unsigned)((N + 31) / 32);
                if (atomicAdd(a.io.gather_counter, 1u) == nwarps - 1u) {
                    *a.io.gather_counter = 0u;
                    __threadfence_system();
                    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.io.gather_flag), "r"(a.io.gather_seq) : "memory");
                }
            }
        }
    };
    auto patch_rows_a4 = [&]() {            // new head -> slots [A, A+12) of my row, new action -> the A slots after it (in place)
        if (live) {
            float* row = xs + (size_t)lane * od;
            float4* r4 = reinterpret_cast<float4*>(row + 4);
            r4[0] = make_float4(h[0], h[1], h[2], h[3]); r4[1] = make_float4(h[4], h[5], h[6], h[7]); r4[2] = make_float4(h[8], h[9], h[10], h[11]);
            *reinterpret_cast<float4*>(row + od) = make_float4(act[0], act[1], act[2], act[3]);
        }
        __syncwarp();
    };
    if (A == 4 && a.early_store) {
        if (!stored) { mbar_wait(bar, 0); store_span(); }
        if (want_fin) {                                                            // terminal observations: head from fin_s, history from the span
            __syncwarp();
            const int c4n = od >> 2;
            float4* fin = reinterpret_cast<float4*>(a.io.final_obs + w0 * od);
            for (unsigned m = fin_rows; m; m &= m - 1) {
                const int r = __ffs(m) - 1;
                const float4 ar = make_float4(__shfl_sync(0xffffffffu, act[0], r), __shfl_sync(0xffffffffu, act[1], r),
                                              __shfl_sync(0xffffffffu, act[2], r), __shfl_sync(0xffffffffu, act[3], r));
                for (int c = lane; c < c4n; c += 32)
                    fin[r * c4n + c] = c < 3 ? reinterpret_cast<const float4*>(fin_s + 12 * r)[c] : (c < c4n - 1 ? shifted[r * c4n + c] : ar);
            }
        }
        QS_STAMP(7);
        // the bulk store wrote stale values into the head and newest-action slots of every row: wait until it has completed,
        // then overwrite them (same addresses: the generic stores must be ordered after the asynchronous ones)
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
        if (live) {
            float4* row = reinterpret_cast<float4*>(a.io.obs + i * od);
            row[0] = make_float4(h[0], h[1], h[2], h[3]); row[1] = make_float4(h[4], h[5], h[6], h[7]); row[2] = make_float4(h[8], h[9], h[10], h[11]);
            row[(od >> 2) - 1] = make_float4(act[0], act[1], act[2], act[3]);
        }
        QS_STAMP(8);
        if (a.io.obs_gather) { patch_rows_a4(); gather(shifted); }      // (the early bulk store has completed: shared memory is free)
        return;
    }
    mbar_wait(bar, 0);
    QS_STAMP(7);
    if (A == 4) {
        patch_rows_a4();
        if (lane == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        store_span();
        if (want_fin) {                                                            // terminal observations: head from fin_s, history from the span
            const int c4n = od >> 2;
            float4* fin = reinterpret_cast<float4*>(a.io.final_obs + w0 * od);
            for (unsigned m = fin_rows; m; m &= m - 1) {
                const int r = __ffs(m) - 1;
                for (int c = lane; c < c4n; c += 32)
                    fin[r * c4n + c] = c < 3 ? reinterpret_cast<const float4*>(fin_s + 12 * r)[c] : shifted[r * c4n + c];
            }
        }
    } else {
        // funnel shift by one float: ys[j] = xs[j + 1], four floats per thread and iteration (LDS.128 + one shuffle + STS.128)
        const int n4 = (rows * od + 3) >> 2;
        for (int j = lane; j < ((n4 + 31) & ~31); j += 32) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j <= n4) v = reinterpret_cast<const float4*>(xs)[j];              // j == n4: the float past the span (padding)
            float nx = __shfl_down_sync(0xffffffffu, v.x, 1);
            if (lane == 31 && j + 1 <= n4) nx = xs[4 * (j + 1)];
            if (j < n4) reinterpret_cast<float4*>(ys)[j] = make_float4(v.y, v.z, v.w, nx);
        }
        __syncwarp();
        if (live) {
            float* row = ys + (size_t)lane * od;                                  // od = 12 + B: odd word stride for B = 15, conflict-free
#pragma unroll
            for (int k = 0; k < 12; ++k) row[k] = h[k];
            row[od - 1] = act[0];
        }
        __syncwarp();
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(span_dst), "r"(smem_u32(ys)), "r"(span_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (want_fin) {
            float* fin = a.io.final_obs + w0 * od;
            for (unsigned m = fin_rows; m; m &= m - 1) {
                const int r = __ffs(m) - 1;
                for (int c = lane; c < od; c += 32) fin[r * od + c] = c < 12 ? fin_s[12 * r + c] : ys[r * od + c];
            }
        }
    }
    QS_STAMP(8);
    if (a.io.obs_gather) { gather(A == 4 ? (const void*)shifted : (const void*)ys); return; }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the bulk store's reads
    QS_STAMP(9);
}

template <int A, bool TASK, bool RESET, bool RPYF, int WARPS>
cudaError_t launch_one(const StepArgs& a, cudaStream_t s) {
    const long long warps = a.n_warps > 0 ? a.n_warps : (a.N + 31) / 32;
    const int blocks = (int)((warps + WARPS - 1) / WARPS);
    const bool fin = RESET && a.io.final_obs != nullptr;
    const size_t sm = (size_t)WARPS * FastSmem<A>::total_bytes(a.obs_dim, fin);
    static const bool pdl = !(getenv("QS_PDL") && atoi(getenv("QS_PDL")) == 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(32 * WARPS); cfg.dynamicSmemBytes = sm; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    if (sm > 48 * 1024)
        cudaFuncSetAttribute(step_fast_kernel<A, TASK, RESET, RPYF, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    return cudaLaunchKernelEx(&cfg, step_fast_kernel<A, TASK, RESET, RPYF, WARPS>, a);
}

template <int A, int WARPS>
cudaError_t launch_modes(const StepArgs& a, cudaStream_t s) {
    const bool task = a.task == QS_TASK_HOVER, reset = a.flags & QS_FLAG_AUTORESET_SAME_STEP, rpyf = a.flags & QS_FLAG_RPY_F32;
    const int key = (task ? 4 : 0) | (reset ? 2 : 0) | (rpyf ? 1 : 0);
    switch (key) {
        case 0: return launch_one<A, false, false, false, WARPS>(a, s);
        case 1: return launch_one<A, false, false, true, WARPS>(a, s);
        case 2: return launch_one<A, false, true, false, WARPS>(a, s);
        case 3: return launch_one<A, false, true, true, WARPS>(a, s);
        case 4: return launch_one<A, true, false, false, WARPS>(a, s);
        case 5: return launch_one<A, true, false, true, WARPS>(a, s);
        case 6: return launch_one<A, true, true, false, WARPS>(a, s);
        default: return launch_one<A, true, true, true, WARPS>(a, s);
    }
}

}  // namespace

// The fast kernels cover: RL observations with an action buffer, act RPM / ONE_D_RPM, no DYN+ effects, D in {1,2,4,...,32},
// autoreset SAME_STEP or none without the opt-in clear flags, a span that is 16-byte aligned and sized for every warp.
bool step_fast_eligible(const StepArgs& a) {
    if (a.act_type != QS_ACT_RPM && a.act_type != QS_ACT_ONE_D_RPM) return false;
    if ((a.effects & 7u) != 0u) return false;
    if (a.flags & ~(unsigned)(QS_FLAG_AUTORESET_SAME_STEP | QS_FLAG_RPY_F32)) return false;
    if (a.D > 32 || (a.D & (a.D - 1)) != 0) return false;
    if (!a.io.obs || !a.io.obs_prev || a.io.act_buffer_size < 1 || a.io.dw_fz) return false;
    if (!aligned16(a.io.obs) || !aligned16(a.io.obs_prev)) return false;
    if (a.io.final_obs && !aligned16(a.io.final_obs)) return false;
    if ((a.N % 32) * (long long)a.obs_dim % 4 != 0) return false;                  // ragged last warp: its span must stay a 16-byte multiple
    if ((a.flags & QS_FLAG_AUTORESET_SAME_STEP) && !a.st.reset_head) return false;
    if ((size_t)FastSmem<4>::total_bytes(a.obs_dim, true) > kStageLimit) return false;      // long action buffers (240 Hz control)
    return true;
}

cudaError_t launch_step_fast(const StepArgs& a_in, cudaStream_t s) {
#ifdef QS_TIMELINE
    StepArgs a = a_in;
    a.dbg_slot = g_dbg_slot;
#else
    const StepArgs& a = a_in;
#endif
    static const int warps = getenv("QS_FAST_WARPS") ? atoi(getenv("QS_FAST_WARPS")) : 1;      // measured default
    if (a.A == 4) {
        if (warps == 4) return launch_modes<4, 4>(a, s);
        if (warps == 2) return launch_modes<4, 2>(a, s);
        return launch_modes<4, 1>(a, s);
    }
    if (warps == 4) return launch_modes<1, 4>(a, s);
    if (warps == 2) return launch_modes<1, 2>(a, s);
    return launch_modes<1, 1>(a, s);
}

}  // namespace qsi

#ifdef QS_TIMELINE
extern "C" int qs_debug_timeline(unsigned long long* host_out, int n_words, int slot) {
    return (int)cudaMemcpyFromSymbol(host_out, g_timeline, (size_t)n_words * 8, (size_t)(slot & 3) * 8192 * 16 * 8);
}
extern "C" void qs_debug_set_slot(int slot) { g_dbg_slot = slot; }
#endif
