// slicer_pipe.cuh -- the pipelined mode of the streaming slicer (included by slicer_fast.cuh).
//
// Same contract as the synchronous tile loop of slicer_fast_kernel (transition_sink.py:55-82, guessed thresholds that
// are proven afterwards), organised as a producer / consumer pipeline over mbarriers instead of two block barriers
// per tile:
//
//  * NW worker warps classify tile k against guessed thresholds (one guess per warp and tile: the warp's R*128
//    contiguous samples are one chunk), keep what the ring slots will hold in registers, publish one record per warp
//    (fixed-point sum of the admitted x - prev, max |x - prev|, the smallest distance of any sample to a guessed
//    threshold) and the tile's class bitmap, and only then wait for the verdict on tile k-1 and write its ring slots.
//    They never wait for tile k's own verdict before starting tile k+1.
//  * The judge warp (lane = chunk) turns the records of tile k into the window sum at every chunk start as an
//    interval, proves every chunk's guess (the band the true thresholds can lie in is strictly inside guess +- margin),
//    carries the interval on, publishes the guesses of tile k+2, and -- as soon as the workers are done reading a
//    stage -- has one lane issue the bulk copy (cp.async.bulk, completion on an mbarrier) of tile k+S into it.
//  * The mapper warp (lane = 128 samples) derives from the bitmap what the hysteresis needs: whether a HIGH sample
//    follows a LOW sample closely enough for cur_state == 2 to matter (then val != class and the tile is not ours),
//    and the carries (val of the last sample, last LOW sample, start of its run).
//  * The ring slots are rewritten while a tile is classified; what they held goes to the tile's stage in place of the
//    samples (an undo log), so that a refused tile -- and the tile classified ahead of its verdict -- can be taken back.
//  * A tile whose guess cannot be proven (a sample too close to a threshold for what is known about the window sum) is
//    taken back and classified again inside the pipeline by the precise pass (every sample against its own guessed
//    window sum: the chunk's measured start + the steps before the sample), which the judge checks the same way.
//  * What the precise pass cannot prove either, and tiles where the hysteresis may matter, end the pipelined run at
//    that tile: the ring is as before it, and the synchronous loop settles it (exact fix-point, exact path) before
//    the pipeline is entered again.
//
// Ordering of the ring: the slots tile k writes are read next by tiles >= k + L/T - 1 >= k + 2 (L >= 3T is required by
// the caller).  A worker starts tile k' after the verdict on k'-2, i.e. after every worker's arrival on rec_full[k'-2],
// i.e. after every ring write of tiles <= k'-2.
#pragma once

namespace nfc {

// ---------------------------------------------------------------- mbarrier / bulk copy (PTX)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(unsigned long long *b) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *b) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *b, unsigned bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(b)), "r"(bytes)
                 : "memory");
}
// Waits for the phase of the given parity to complete.  A wait that does not end (a protocol error: each try_wait already
// suspends the thread for a while) traps instead of hanging the device.
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .u32 n;\n\t"
        "mov.u32 n, 0;\n\t"
        "PIPE_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
        "@p bra PIPE_DONE;\n\t"
        "add.u32 n, n, 1;\n\t"
        "setp.lt.u32 p, n, 0x100000;\n\t"
        "@p bra PIPE_WAIT;\n\t"
        "trap;\n\t"
        "PIPE_DONE:\n\t"
        "}" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (the TMA unit, no tensor map), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- packed float helpers (no `volatile`: free to schedule)
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// min(m, |a|, |b|), NaN if any operand is
__device__ __forceinline__ float min3nan_abs(float m, float a, float b) {
    float r;
    asm("{\n\t.reg .f32 aa, bb;\n\tabs.f32 aa, %2;\n\tabs.f32 bb, %3;\n\tmin.NaN.f32 %0, %1, aa, bb;\n\t}" : "=f"(r) : "f"(m), "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float max3_abs(float m, float a, float b) {
    float r;
    asm("{\n\t.reg .f32 aa, bb;\n\tabs.f32 aa, %2;\n\tabs.f32 bb, %3;\n\tmax.f32 %0, %1, aa, bb;\n\t}" : "=f"(r) : "f"(m), "f"(a), "f"(b));
    return r;
}

// ---------------------------------------------------------------- shared state of the pipelined mode
struct __align__(16) PipeRec {  // one warp's chunk (R * 128 samples) of one tile
    int S;       // round(sum of admitted (x - prev) / q), summed over the lanes (and rows)
    float amax;  // cheap pass: max |x - prev| over the admitted samples; precise pass: sum of |x - prev|, in steps, rounded up
    float m;     // cheap pass: min over the samples of the distance to the nearer guessed threshold; precise pass: the smallest
                 // slack of any lane, in window-sum units (NaN: not usable)
    int pad;
};

enum { PIPE_CMD_ENTER = 1, PIPE_CMD_QUIT = 2 };
enum { PV_ACCEPT = 0, PV_ABORT = 1, PV_REDO = 2 };
static const int PIPE_BAR_RUN = 2, PIPE_BAR_PARK = 3;  // named barriers over all threads of the CTA (workers + judge + mapper)
template <int ID, int N>
__device__ __forceinline__ void named_bar_sync() {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory");
}

template <int NW, int R, int S>
struct __align__(16) PipeShared {
    unsigned long long x_full[S];    // stage s holds the samples of tile k, k % S == s
    unsigned long long rec_full[2];  // all workers have published tile k (k & 1) -- its cheap pass, or its precise pass
    unsigned long long verdict[2];   // judge and mapper have judged tile k (k & 1); the guesses of tile k+2 are published
    float4 G[4][NW];                 // per warp of tile k (k & 3): -centre and radius of the guessed thresholds, 1/q
    float4 J[4][NW];                 // the judge's copy of the guesses of tile k (k & 3): offset of the guessed window sum from the
                                     // one it was made from, guessed LOW threshold, HIGH threshold * 2^-19, fixed-point step
    int gok[4];                      // the guesses of tile k (k & 3) are usable
    float tlm[2], thm[2];            // thresholds at the window sum of tile k's first sample (k & 1), for its precise pass
    PipeRec recs[2][NW];
    uint32_t bm[2][NW * R * 8];      // the tile's bitmap words, chunk of 128 samples major (as FastShared::bm)
    int vjudge[2], vjudge2[2];       // the judge's verdict (PV_*) on the cheap pass / on the precise pass of tile k (k & 1)
    int vst2[2], vst2b[2];           // the mapper's: the hysteresis may matter
    // what the precise pass of a refused tile assumes (judge -> workers)
    float redo_c0[NW];               // measured window sum at each chunk's first sample, less the tile's
    float redo_q, redo_invq;
    float hw_out, pad_f;             // the run's results: half width of the window sum's interval (judge), its middle (mapper)
    double ssm_out;
    int cmd, t0, K, done;            // command to the parked warps; first tile and number of tiles of the run; tiles proven
    int inited, redone, pad_[2];
};

// A stage: the tile's samples, overwritten in place by the undo log (what the ring slots held) while the tile is classified.
// 16-bit samples are narrower than the log's floats: their log follows the samples instead.
template <int NW, int R, int ITEM>
struct PipeStage {
    static const int T = NW * R * FAST_CH;
    static const int log_ofs = ITEM == 4 ? 0 : T * ITEM;
    static const int bytes = log_ofs + T * 4;
};

template <int NW, int R>
struct PipeConsts {
    static const int NC = NW * R;            // chunks of 128 samples per tile
    static const int CHS = R * FAST_CH;      // samples per warp and tile
    static const int T = NW * CHS;           // samples per tile
};

// ---------------------------------------------------------------- workers
// Two samples against the guessed thresholds centre -+ radius (ncg2 = {-centre, -centre}).
struct PipeAcc {
    unsigned long long s2;  // two running sums of n - prev
    float amax, m;
};
__device__ __forceinline__ void pipe_pair(float x0, float x1, float p0, float p1, unsigned long long ncg2, float rg, float nrg, PipeAcc &a,
                                          unsigned &NL0, unsigned &NL1, unsigned &H0, unsigned &H1, float &n0, float &n1) {
    float u0, u1;
    unpack2(add2(pack2(x0, x1), ncg2), u0, u1);
    const bool pnl0 = u0 > nrg, ph0 = u0 > rg, pnl1 = u1 > nrg, ph1 = u1 > rg;
    NL0 = __ballot_sync(FULL, pnl0);
    H0 = __ballot_sync(FULL, ph0);
    NL1 = __ballot_sync(FULL, pnl1);
    H1 = __ballot_sync(FULL, ph1);
    n0 = (pnl0 && !ph0) ? x0 : p0;  // transition_sink.py:75-81: only the MID branch admits the sample
    n1 = (pnl1 && !ph1) ? x1 : p1;
    a.m = min3nan_abs(a.m, fabsf(u0) - rg, fabsf(u1) - rg);
    const unsigned long long dd = sub2(pack2(n0, n1), pack2(p0, p1));
    a.s2 = add2(a.s2, dd);
    float d0, d1;
    unpack2(dd, d0, d1);
    a.amax = max3_abs(a.amax, d0, d1);
}

// One row (128 samples) of the precise pass (as fast_row<true> of the synchronous loop): every sample against its own
// guessed window sum, c0g (the row's first sample, less the tile's) + the steps of the lanes before it + the steps before
// it inside the lane.  TLb / THb: thresholds at the window sum of the tile's first sample.  Returns the row's fixed-point
// sum of n - prev, of |n - prev| (rounded up), and the smallest slack of any lane in window-sum units (NaN: not usable).
__device__ __forceinline__ void pipe_row_precise(const float4 x4, const float4 pv4, const float c0g, const float TLb, const float THb,
                                                 const float loLf, const float hiLf, const float invLo, const float invHi, const float invq,
                                                 const int lane, float4 &n4, unsigned (&NLm)[4], unsigned (&Hm)[4], int &S, int &A,
                                                 float &slack) {
    const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
    const float ps[4] = {pv4.x, pv4.y, pv4.z, pv4.w};
    // first the steps under the classes the thresholds at the row's first sample give
    const float thL = fmaf(c0g, loLf, TLb), thH = fmaf(c0g, hiLf, THb);
    float s1 = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const bool adm = xs[j] > thL && !(xs[j] > thH);
        s1 += adm ? xs[j] - ps[j] : 0.0f;
    }
    float incA = s1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(FULL, incA, o);
        if (lane >= o) incA += v;
    }
    const float PA = incA - s1;
    // ... then every sample against its own guessed window sum, in order inside the lane; margins in window-sum units
    const float base = c0g + PA;
    const float tl0 = fmaf(base, loLf, TLb), th0 = fmaf(base, hiLf, THb);
    float run = 0.0f, wmin = INFINITY, a = 0.0f;
    float nn[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float tl = fmaf(run, loLf, tl0), th = fmaf(run, hiLf, th0);
        const bool pnl = xs[j] > tl, ph = xs[j] > th;
        NLm[j] = __ballot_sync(FULL, pnl);
        Hm[j] = __ballot_sync(FULL, ph);
        nn[j] = (pnl && !ph) ? xs[j] : ps[j];
        const float d = nn[j] - ps[j];
        run += d;
        a += fabsf(d);
        wmin = fmin_nan(wmin, fmin_nan(fabsf(xs[j] - tl) * invLo, fabsf(xs[j] - th) * invHi));
    }
    n4 = make_float4(nn[0], nn[1], nn[2], nn[3]);
    float incB = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(FULL, incB, o);
        if (lane >= o) incB += v;
    }
    // slack of the lane: its margin less what the second classification moved the steps before it by
    float mL = wmin - fabsf((incB - run) - PA) * 1.001f;
    const float af = a * (invq * (1.0f + 0x1p-20f));
    if (!(af < 4194304.0f)) mL = __int_as_float(0x7fc00000);  // the lane's sums do not fit 2^22 steps (or are NaN)
    const int si = __float2int_rn(run * invq);
    const int ai = __float2int_ru(fminf(af, 4194304.0f));
    S = __reduce_add_sync(FULL, si);
    A = __reduce_add_sync(FULL, ai);
    slack = redux_min_nan(mL);
}

// Shared-memory accesses by 32-bit address in the explicit state space: the hot loop carries no generic pointers, and
// constant offsets fold into the instructions.  All `volatile`: they keep their program order.
template <int OFF>
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ short4 lds_s4(uint32_t a) {
    short4 v;
    asm volatile("ld.shared.v4.s16 {%0, %1, %2, %3}, [%4+%5];" : "=h"(v.x), "=h"(v.y), "=h"(v.z), "=h"(v.w) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts_f4(uint32_t a, const float4 v) {
    asm volatile("st.shared.v4.f32 [%0+%1], {%2, %3, %4, %5};" ::"r"(a), "n"(OFF), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <int OFF>
__device__ __forceinline__ void sts_u4(uint32_t a, unsigned x, unsigned y, unsigned z, unsigned w) {
    asm volatile("st.shared.v4.u32 [%0+%1], {%2, %3, %4, %5};" ::"r"(a), "n"(OFF), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ unsigned lds_u32(uint32_t a) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t addr, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .u32 n;\n\t"
        "mov.u32 n, 0;\n\t"
        "PIPE_WAIT_A:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
        "@p bra PIPE_DONE_A;\n\t"
        "add.u32 n, n, 1;\n\t"
        "setp.lt.u32 p, n, 0x100000;\n\t"
        "@p bra PIPE_WAIT_A;\n\t"
        "trap;\n\t"
        "PIPE_DONE_A:\n\t"
        "}" ::"r"(addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t addr) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(addr) : "memory");
}

// The precise pass over tile j of the run (cold path, kept out of the hot loop's code): the ring slots of the tile were
// taken back; the samples come from global memory again (the stage holds the log).  Leaves the warp's record in registers.
template <int NW, int R, int S, int KIND>
__device__ __noinline__ void pipe_worker_redo(PipeShared<NW, R, S> &ps, float *ring, char *lrow, const char *src, const FastPlan &plan,
                                              const int L, const float pcm_scale, const int bp, const int warp, const int lane, int slot,
                                              int &Stot_out, float &Atot_out, float &mslack_out) {
    const float TLb = ps.tlm[bp], THb = ps.thm[bp], rq = ps.redo_q, rinvq = ps.redo_invq;
    float c0g = ps.redo_c0[warp];
    uint32_t *bms = &ps.bm[bp][warp * (R * 8)];
    int Stot = 0, Atot = 0;
    float mslack = INFINITY;
    float4 xv[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        if (KIND == IN_PCM_S16) {
            const short4 sv = __ldg(reinterpret_cast<const short4 *>(src + r * (FAST_CH * 2)));
            xv[r] = make_float4(env_real(__fdiv_rn((float)sv.x, pcm_scale)), env_real(__fdiv_rn((float)sv.y, pcm_scale)),
                                env_real(__fdiv_rn((float)sv.z, pcm_scale)), env_real(__fdiv_rn((float)sv.w, pcm_scale)));
        } else {
            xv[r] = ldg_stream4(reinterpret_cast<const float4 *>(src + r * (FAST_CH * 4)));
            if (KIND == IN_REAL_F32) {
                xv[r].x = env_real(xv[r].x); xv[r].y = env_real(xv[r].y); xv[r].z = env_real(xv[r].z); xv[r].w = env_real(xv[r].w);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        const float4 p4 = *reinterpret_cast<const float4 *>(ring + slot);
        unsigned NL[4], H[4];
        float4 n4;
        int Sr, Ar;
        float mr;
        pipe_row_precise(xv[r], p4, c0g, TLb, THb, plan.loLf, plan.hiLf, plan.invLo, plan.invHi, rinvq, lane, n4, NL, H, Sr, Ar, mr);
        *reinterpret_cast<float4 *>(ring + slot) = n4;
        *reinterpret_cast<float4 *>(lrow + r * (FAST_CH * 4)) = p4;
        if (lane == 0) {
            uint4 *bw = reinterpret_cast<uint4 *>(bms + r * 8);
            bw[0] = make_uint4(NL[0], NL[1], NL[2], NL[3]);
            bw[1] = make_uint4(H[0], H[1], H[2], H[3]);
        }
        Stot += Sr;
        Atot += Ar;
        mslack = fmin_nan(mslack, mr);
        c0g += (float)Sr * rq;  // the next row starts where the judge's sums will put it
        slot += FAST_CH;
        if (slot >= L) slot -= L;
    }
    Stot_out = Stot;
    Atot_out = (float)Atot;
    mslack_out = mslack;
}

// Returns the number of tiles of the run that were proven (K when the whole run was); the ring holds exactly those.
template <int NW, int R, int S, int KIND>
__device__ __noinline__ int pipe_worker(PipeShared<NW, R, S> &ps, float *ring, char *stage0, const FastPlan &plan, const int L,
                                        const float pcm_scale, const int warp, const int lane) {
    typedef PipeConsts<NW, R> C;
    constexpr int ITEM = KIND == IN_PCM_S16 ? 2 : 4;
    constexpr int stage_bytes = PipeStage<NW, R, ITEM>::bytes;
    constexpr int log_ofs = PipeStage<NW, R, ITEM>::log_ofs;
    constexpr int tile_bytes = C::T * ITEM;
    constexpr int XROW = FAST_CH * ITEM, LROW = FAST_CH * 4;  // bytes between a thread's rows: samples, ring / log
    static_assert(R == 4 || R == 2, "rows are unrolled by hand below");
    const int t0 = ps.t0, K = ps.K;
    const int xoff = (warp * C::CHS + lane * 4) * ITEM;  // this thread's first sample inside a tile of the input
    // shared-memory addresses of this thread's / this warp's parts
    const uint32_t a_ring = smem_u32(ring), a_ring_end = a_ring + (uint32_t)L * 4u;
    const uint32_t a_x = smem_u32(stage0) + (uint32_t)xoff;                                           // samples inside stage 0
    const uint32_t a_log = smem_u32(stage0) + (uint32_t)(log_ofs + (warp * C::CHS + lane * 4) * 4);  // undo log inside stage 0
    const uint32_t a_G = smem_u32(&ps.G[0][warp]);      // + (k & 3) * NW * 16
    const uint32_t a_rec = smem_u32(&ps.recs[0][warp]);  // + b * NW * 16
    const uint32_t a_bm = smem_u32(&ps.bm[0][warp * (R * 8)]);  // + b * NW * R * 32
    const uint32_t a_xfull = smem_u32(&ps.x_full[0]), a_recfull = smem_u32(&ps.rec_full[0]), a_verdict = smem_u32(&ps.verdict[0]);
    const uint32_t a_vj = smem_u32(&ps.vjudge[0]), a_vs = smem_u32(&ps.vst2[0]);  // vjudge2 / vst2b follow 8 bytes behind each
    typedef PipeShared<NW, R, S> PS;
    static_assert(offsetof(PS, vjudge2) == offsetof(PS, vjudge) + 8 && offsetof(PS, vst2b) == offsetof(PS, vst2) + 8,
                  "second verdicts follow the first");
    uint32_t a_slot = a_ring + 4u * (uint32_t)((plan.tile0_pos + (int64_t)t0 * C::T + (int64_t)warp * C::CHS + (int64_t)lane * 4) % L);
    const uint32_t slot_step = 4u * (uint32_t)(C::T % L);
    uint32_t *bm_k = plan.bm_base + (size_t)t0 * (C::NC * 8) + warp * (R * 8) + lane;  // this lane's bitmap word of tile k
    const int emit_from = plan.t_emit - t0;  // tiles of the run from this one on are written to the bitmap
    const char *src0 = plan.xbase + (int64_t)t0 * tile_bytes + xoff;

    auto row_addr = [&](uint32_t first, int r) -> uint32_t {  // ring address of this thread's row r of a tile whose row 0 is at `first`
        uint32_t a = first + (uint32_t)(r * LROW);
        if (a >= a_ring_end) a -= (uint32_t)L * 4u;
        return a;
    };
    auto undo = [&](int stage, uint32_t first) {  // the ring slots of a tile as they were before it
        const uint32_t lg = a_log + (uint32_t)(stage * stage_bytes);
        sts_f4<0>(row_addr(first, 0), lds_f4<0>(lg));
        sts_f4<0>(row_addr(first, 1), lds_f4<LROW>(lg));
        if (R == 4) {
            sts_f4<0>(row_addr(first, 2), lds_f4<2 * LROW>(lg));
            sts_f4<0>(row_addr(first, 3), lds_f4<3 * LROW>(lg));
        }
    };
    unsigned xph = 0u, vph = 0u;  // bit s / bit b: the parity of the phase to wait for next on x_full[s] / verdict[b]
    // the verdict on the tile with parity b (the cheap pass's, or the precise pass's)
    auto wait_verdict = [&](int b, int second) -> int {
        mbar_wait_a(a_verdict + 8u * b, (vph >> b) & 1u);
        vph ^= 1u << b;
        const unsigned vs = lds_u32(a_vs + 4u * b + 8u * second), vj = lds_u32(a_vj + 4u * b + 8u * second);
        return vs ? (int)PV_ABORT : (int)vj;
    };
    // record, bitmap words and the arrival of tile k (bm_t: this lane's bitmap word of that tile in global memory)
    auto publish = [&](bool emit, uint32_t *bm_t, int b, int Ssum, float amax, float m) {
        if (lane == 0) sts_u4<0>(a_rec + (uint32_t)(b * NW * 16), (unsigned)Ssum, __float_as_uint(amax), __float_as_uint(m), 0u);
        __syncwarp();
        if (emit && lane < R * 8) *bm_t = lds_u32(a_bm + (uint32_t)(b * NW * R * 32 + lane * 4));
        if (lane == 0) mbar_arrive_a(a_recfull + 8u * b);
    };

    uint32_t a_slot_prev = 0u;
    int st_prev = 0, st = 0, k = 0;
    bool pending = false;  // tile k-1 awaits its verdict
#pragma unroll 1
    for (;;) {
        const int b = k & 1;
        int Ssum = 0;
        unsigned amax_u = 0u;
        float m = 0.0f;
        if (k < K) {
            // ------------------------------------------------------------ tile k against its guesses
            mbar_wait_a(a_xfull + 8u * st, (xph >> st) & 1u);
            xph ^= 1u << st;
            const float4 g = lds_f4<0>(a_G + (uint32_t)((k & 3) * NW * 16));
            const unsigned long long ncg2 = pack2(g.x, g.x);
            const float rg = g.y, nrg = -g.y, invq = g.z;
            const uint32_t ax = a_x + (uint32_t)(st * stage_bytes), alog = a_log + (uint32_t)(st * stage_bytes);
            const uint32_t abm = a_bm + (uint32_t)(b * NW * R * 32);
            uint32_t ar[R];
#pragma unroll
            for (int r = 0; r < R; r++) ar[r] = row_addr(a_slot, r);
            // all loads of the tile first: the rows' dependency chains overlap
            float4 xv[R], pv[R];
            if (KIND == IN_PCM_S16) {
                short4 sv[R];
                sv[0] = lds_s4<0>(ax);
                sv[1] = lds_s4<XROW>(ax);
                if (R == 4) { sv[2] = lds_s4<2 * XROW>(ax); sv[3] = lds_s4<3 * XROW>(ax); }
#pragma unroll
                for (int r = 0; r < R; r++)
                    xv[r] = make_float4(env_real(__fdiv_rn((float)sv[r].x, pcm_scale)), env_real(__fdiv_rn((float)sv[r].y, pcm_scale)),
                                        env_real(__fdiv_rn((float)sv[r].z, pcm_scale)), env_real(__fdiv_rn((float)sv[r].w, pcm_scale)));
            } else {
                xv[0] = lds_f4<0>(ax);
                xv[1] = lds_f4<XROW>(ax);
                if (R == 4) { xv[2] = lds_f4<2 * XROW>(ax); xv[3] = lds_f4<3 * XROW>(ax); }
            }
#pragma unroll
            for (int r = 0; r < R; r++) pv[r] = lds_f4<0>(ar[r]);
            PipeAcc a;
            a.s2 = 0ull;
            a.amax = 0.0f;
            a.m = INFINITY;
            unsigned NL[R][4], H[R][4];
            float4 n4[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                float4 x4 = xv[r];
                if (KIND == IN_REAL_F32) { x4.x = env_real(x4.x); x4.y = env_real(x4.y); x4.z = env_real(x4.z); x4.w = env_real(x4.w); }
                pipe_pair(x4.x, x4.y, pv[r].x, pv[r].y, ncg2, rg, nrg, a, NL[r][0], NL[r][1], H[r][0], H[r][1], n4[r].x, n4[r].y);
                pipe_pair(x4.z, x4.w, pv[r].z, pv[r].w, ncg2, rg, nrg, a, NL[r][2], NL[r][3], H[r][2], H[r][3], n4[r].z, n4[r].w);
                sts_f4<0>(ar[r], n4[r]);
                if (r == 0) sts_f4<0>(alog, pv[r]);
                if (r == 1) sts_f4<LROW>(alog, pv[r]);
                if (r == 2) sts_f4<2 * LROW>(alog, pv[r]);
                if (r == 3) sts_f4<3 * LROW>(alog, pv[r]);
                if (lane == 0) {
                    if (r == 0) { sts_u4<0>(abm, NL[r][0], NL[r][1], NL[r][2], NL[r][3]); sts_u4<16>(abm, H[r][0], H[r][1], H[r][2], H[r][3]); }
                    if (r == 1) { sts_u4<32>(abm, NL[r][0], NL[r][1], NL[r][2], NL[r][3]); sts_u4<48>(abm, H[r][0], H[r][1], H[r][2], H[r][3]); }
                    if (r == 2) { sts_u4<64>(abm, NL[r][0], NL[r][1], NL[r][2], NL[r][3]); sts_u4<80>(abm, H[r][0], H[r][1], H[r][2], H[r][3]); }
                    if (r == 3) { sts_u4<96>(abm, NL[r][0], NL[r][1], NL[r][2], NL[r][3]); sts_u4<112>(abm, H[r][0], H[r][1], H[r][2], H[r][3]); }
                }
            }
            // the warp's record
            float sa, sb;
            unpack2(a.s2, sa, sb);
            const float ssum = sa + sb;
            m = a.m;
            if (!(fabsf(ssum) * invq < 4194304.0f)) m = __int_as_float(0x7fc00000);  // the lane's sum does not fit 2^22 steps (or is NaN)
            const int si = __float2int_rn(ssum * invq);
            Ssum = __reduce_add_sync(FULL, si);
            amax_u = __reduce_max_sync(FULL, __float_as_uint(a.amax));  // amax >= 0: ordered like its bit pattern
            m = redux_min_nan(m);
        }
        if (pending) {
            // ------------------------------------------------------------ the verdict on tile k-1
            // (no tile is announced two ahead of an open verdict: ring ordering, depth of the undo logs)
            const int bp = b ^ 1;
            int v = wait_verdict(bp, 0);
            if (v != PV_ACCEPT) {
                if (k < K) undo(st, a_slot);
                undo(st_prev, a_slot_prev);
                if (v != PV_REDO) return k - 1;
                // ---- the precise pass over tile k-1
                int Stot;
                float Atot, mslack;
                pipe_worker_redo<NW, R, S, KIND>(ps, ring, stage0 + log_ofs + (warp * C::CHS + lane * 4) * 4 + st_prev * stage_bytes,
                                                 src0 + (int64_t)(k - 1) * tile_bytes, plan, L, pcm_scale, bp, warp, lane,
                                                 (int)((a_slot_prev - a_ring) >> 2), Stot, Atot, mslack);
                publish(k - 1 >= emit_from, bm_k - (C::NC * 8), bp, Stot, Atot, mslack);
                v = wait_verdict(bp, 1);
                if (v != PV_ACCEPT) {
                    undo(st_prev, a_slot_prev);
                    return k - 1;
                }
                pending = false;
                if (k >= K) return K;
                continue;  // tile k again: the mapper has its samples copied again and new guesses made
            }
            pending = false;
        }
        if (k >= K) return K;
        publish(k >= emit_from, bm_k, b, Ssum, __uint_as_float(amax_u), m);
        pending = true;
        a_slot_prev = a_slot;
        st_prev = st;
        a_slot += slot_step;
        if (a_slot >= a_ring_end) a_slot -= (uint32_t)L * 4u;
        bm_k += C::NC * 8;
        if (++st == S) st = 0;
        k++;
    }
}

// ---------------------------------------------------------------- judge and mapper
// fixed-point step: a power of two near a_est * 2^-27.  A lane's sum must stay below 2^22 steps (a_est / 32, a_est being an
// upper bound of the tile's sum of |x - prev| of late: 128 times the lane's share), so that the sums of a tile's chunks fit int32.
__device__ __forceinline__ bool pipe_step(float a_est, float &q, float &invq) {
    const unsigned ae = (__float_as_uint(a_est) >> 23) & 0xffu;
    const bool ok = ae > 45u && ae < 250u;
    const unsigned qe = ok ? ae - 27u : 127u;
    q = __uint_as_float(qe << 23);
    invq = __uint_as_float((254u - qe) << 23);
    return ok;
}

// Two warps share what has to happen per tile beside the workers; each is a serial chain of dependent instructions, and
// the longer of the two bounds how fast the workers may go:
//
//  * the judge only proves (or refuses) the guesses of tile k from the workers' records: float / int32 all the way;
//  * the mapper derives the hysteresis maps from the bitmap, keeps the window sum (double), makes the guesses of tile k+2
//    and has the samples of tile k+S copied.
//
// Window sum at the tile's first sample: |true - ssm| <= hw (ssm double, kept by the mapper; hw float rounded up, kept by the
// judge).  The workers start tile j+2 without waiting for anything after the verdict on tile j, so its guesses must be
// out by then: they are made after the verdict on tile j-1, from ssm at the start of tile j (ssm_j) and the drift of tile
// j-1 (dr): the guessed window sum of chunk c is gss = ssm_j + goff[c], goff[c] = dr * (2 + (c + 1/2) / NW); at tile j+2's
// start ssm - ssm_j is the drift of tiles j and j+1 (d2 + d1), so (window sum at chunk c's first sample) - gss =
// (d1 + d2) - goff[c] + c0[c], c0 = exclusive prefix of the chunk sums.  The first three tiles of a run (and the three
// after a precise pass) are guessed from the window sum known then.  Guesses live in rings of four tiles (PipeShared::G
// for the workers, ::J the judge's copy).
template <int NW, int R, int S>
__device__ __noinline__ void pipe_judge(PipeShared<NW, R, S> &ps, const FastUni &uni, const FastPlan &plan, const int lane, const int allow_redo) {
    typedef PipeConsts<NW, R> C;
    const int K = ps.K;
    const double ss_lo0 = uni.ss_lo, ss_hi0 = uni.ss_hi;
    const double ssm0 = 0.5 * (ss_lo0 + ss_hi0);
    float hw = (__double2float_ru(__dsub_ru(ss_hi0, ssm0)) + __double2float_ru(__dsub_ru(ssm0, ss_lo0))) * 1.0001f;
    const float loLf = plan.loLf, hiLs = plan.hiLf * (1.0f + 0x1p-20f);
    const bool act = lane < NW;
    float d1 = 0.0f, d2 = 0.0f;  // drifts of the two tiles before the one being judged, as far as its guesses did not know them
    named_bar_sync<PIPE_BAR_RUN, (NW + 2) * 32>();  // barriers, guesses and flags are set up (by the mapper)

    unsigned rph = 0u, vph = 0u;  // bit b: the parity of the phase to wait for next on rec_full[b] / verdict[b]
    int k = 0;
#pragma unroll 1
    for (; k < K; k++) {
        const int b = k & 1;
        const float4 gj = ps.J[k & 3][act ? lane : 0];  // the tile's guesses: goff, guessed LOW threshold, HIGH * 2^-19, step
        const bool okb = ps.gok[k & 3] != 0;
        const float goff = gj.x, gTL = gj.y, gTHs = gj.z, q = gj.w;
        mbar_wait(&ps.rec_full[b], (rph >> b) & 1u);
        rph ^= 1u << b;
        const PipeRec rc = ps.recs[b][act ? lane : 0];
        const int Si = act ? rc.S : 0;
        const float amax = act ? rc.amax : 0.0f;
        int inc = Si;
#pragma unroll
        for (int o = 1; o < NW; o <<= 1) {
            const int v = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += v;
        }
        const int c0i = inc - Si;                            // chunk sums below 2^27 steps each (lane sums below 2^22): no overflow
        int toti = __shfl_sync(FULL, inc, NW - 1);
        const float c0f = (float)c0i * q, Sf = (float)Si * q;
        // upper bounds of the sums of |x - prev|: of the chunk, of the tile
        const float Ahat = amax * ((float)C::CHS * 1.0001f);
        float totA = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(amax))) * ((float)C::T * 1.0002f);
        // error of the measured sums: conversion (half a step per lane and chunk), float rounding (2^-19 of |x - prev|)
        float Ef = fmaf(totA, 0x1p-19f, (float)(16 * NW) * q) * 1.01f;
        const float dsince = d1 + d2;  // ssm - (the window sum the guesses were made from)
        const float slack = (hw + Ef + (fabsf(d1) + fabsf(d2) + fabsf(goff) + fabsf(c0f)) * 0x1p-21f) * 1.001f;
        // inside the chunk the window sum moves within [V, U] of its start: the sums of the negative / positive steps
        const float U = fmaxf(0.5f * (Ahat + Sf), 0.0f) * (1.0f + 0x1p-20f), V = fminf(-0.5f * (Ahat - Sf), 0.0f) * (1.0f + 0x1p-20f);
        const float off = (dsince - goff) + c0f;
        const float dev = fmaxf(fabsf(off + (U + slack)), fabsf(off + (V - slack))) * (1.0f + 0x1p-20f);  // |window sum - guessed| at any sample
        // the guess is proven when no sample lies between it and any value the true threshold can take
        const float need = fmaf(dev, hiLs, gTHs);
        const bool fine = !act || ((rc.m > need) && (gTL - need > 0.0f));
        const bool accept = __all_sync(FULL, fine) && okb;
        // a refused tile goes through the precise pass if its records are numbers (no sample is NaN, the sums fit the fixed point)
        const bool redo = !accept && allow_redo && okb && __all_sync(FULL, !act || rc.m == rc.m);
        if (redo) {
            if (act) ps.redo_c0[lane] = c0f;
            if (lane == 0) {
                ps.redo_q = q;
                ps.redo_invq = __uint_as_float((254u << 23) - __float_as_uint(q));  // q is a power of two
            }
            __syncwarp();
        }
        if (lane == 0) {
            ps.vjudge[b] = accept ? PV_ACCEPT : (redo ? PV_REDO : PV_ABORT);
            mbar_arrive(&ps.verdict[b]);
        }
        float dr = (float)toti * q;
        const float gTH_k = gTHs * 0x1p19f;
        mbar_wait(&ps.verdict[b], (vph >> b) & 1u);  // the mapper's say
        vph ^= 1u << b;
        {
            const volatile int *vs = ps.vst2;
            if (vs[b] || !(accept || redo)) break;
        }
        if (!accept) {
            // ------------------------------------------------------------ the precise pass's records
            mbar_wait(&ps.rec_full[b], (rph >> b) & 1u);
            rph ^= 1u << b;
            const PipeRec r2 = ps.recs[b][act ? lane : 0];
            const float TLm = *(volatile float *)&ps.tlm[b];  // threshold at the window sum of the tile's first sample
            const int S2 = act ? r2.S : 0;
            const float A2 = act ? r2.amax : 0.0f;  // sum of |x - prev| in steps
            int inc2 = S2;
            float incA2 = A2;
#pragma unroll
            for (int o = 1; o < NW; o <<= 1) {
                const int v = __shfl_up_sync(FULL, inc2, o);
                const float va = __shfl_xor_sync(FULL, incA2, o);
                if (lane >= o) inc2 += v;
                incA2 += va;
            }
            const float c0n = (float)(inc2 - S2) * q;  // the window sum at the chunk's first sample (less the tile's) as it is now
            toti = __shfl_sync(FULL, inc2, NW - 1);
            totA = __shfl_sync(FULL, incA2, 0) * q * 1.0002f;
            // error of the sums: conversion (half a step per lane and row), float rounding
            Ef = fmaf(totA, 0x1p-19f, (float)(16 * NW * R) * q) * 1.01f;
            const float slack2 = (hw + Ef + fabsf(c0n) * 0x1p-21f) * 1.001f;
            // the lanes' slack must cover what the chunk's start is off the assumed one by, the interval, and the rounding
            // of the per-sample thresholds
            const float need2 = (fabsf(c0n - c0f) + slack2) * (1.0f + 0x1p-18f) + TLm * (0x1p-21f / loLf);
            const bool fine2 = !act || (r2.m > need2);
            const bool accept2 = __all_sync(FULL, fine2) && TLm > 0.0f;
            if (lane == 0) {
                ps.vjudge2[b] = accept2 ? PV_ACCEPT : PV_ABORT;
                mbar_arrive(&ps.verdict[b]);
            }
            mbar_wait(&ps.verdict[b], (vph >> b) & 1u);  // the mapper's second say
            vph ^= 1u << b;
            {
                const volatile int *vs = ps.vst2b;
                if (vs[b] || !accept2) break;
            }
            d1 = 0.0f;  // the coming tiles are guessed from the window sum after this tile
            d2 = 0.0f;
        } else {
            d2 = d1;
            d1 = dr;
        }
        // ---------------------------------------------------------------- tile k stands: the interval's half width after it
        // (rounding of ssm: below 2^-52 of the window sum, the guessed HIGH threshold is at least 2^-20 of it for windows < 2^20 samples)
        hw = hw + (Ef + gTH_k * 0x1p-30f + hw * 0x1p-22f) * 1.001f;
    }
    if (lane == 0) {
        ps.hw_out = hw;
        ps.done = k;
    }
}

template <int NW, int R, int S, int ITEM>
__device__ __noinline__ void pipe_mapper(PipeShared<NW, R, S> &ps, FastUni &uni, SegCarry &cs, const FastPlan &plan, const double loL,
                                         const double hiL, char *stage0, const int mx, const int lane) {
    typedef PipeConsts<NW, R> C;
    static_assert(C::NC <= 32, "one lane per chunk of 128 samples");
    constexpr unsigned stage_bytes = PipeStage<NW, R, ITEM>::bytes, tile_bytes = C::T * ITEM;
    const int t0 = ps.t0, K = ps.K;
    const char *src0 = plan.xbase + (int64_t)t0 * tile_bytes;
    unsigned xused = 0u, xlast = 0u;  // bit s: a copy was issued into stage s; the parity of the phase the latest one completes
    auto issue = [&](int tile, int stage) {  // the samples of `tile` into `stage` (all lanes call, one issues)
        if (lane == 0) {
            mbar_expect_tx(&ps.x_full[stage], tile_bytes);
            bulk_g2s(stage0 + (size_t)stage * stage_bytes, src0 + (size_t)tile * tile_bytes, tile_bytes, &ps.x_full[stage]);
        }
        if (xused & (1u << stage)) xlast ^= 1u << stage;
        xused |= 1u << stage;
    };
    if (lane == 0) {
        if (ps.inited) {  // barriers of the previous run: nobody waits on them any more
            for (int s = 0; s < S; s++) mbar_inval(&ps.x_full[s]);
            for (int b = 0; b < 2; b++) {
                mbar_inval(&ps.rec_full[b]);
                mbar_inval(&ps.verdict[b]);
            }
        }
        ps.inited = 1;
        for (int s = 0; s < S; s++) mbar_init(&ps.x_full[s], 1);
        for (int b = 0; b < 2; b++) {
            mbar_init(&ps.rec_full[b], NW);
            mbar_init(&ps.verdict[b], 2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    int issued = 0;
    for (; issued < S && issued < K; issued++) issue(issued, issued);
    // ---- state handed over by the synchronous loop
    const double ss_lo0 = uni.ss_lo, ss_hi0 = uni.ss_hi;
    double ssm = 0.5 * (ss_lo0 + ss_hi0);
    const float hw0 = (__double2float_ru(__dsub_ru(ss_hi0, ssm)) + __double2float_ru(__dsub_ru(ssm, ss_lo0))) * 1.0001f;
    float drift = uni.tot_prev, a_est = uni.a_est;
    if (!(a_est > 0.0f)) a_est = __double2float_rd(ss_lo0) * 0x1p-7f;
    const float loLf = plan.loLf, hiLf = plan.hiLf;
    const bool actw = lane < NW;
    const float fc = ((float)lane + 0.5f) * (1.0f / NW);  // the chunk's middle, in tiles
    float thr_min = 3.0e38f, thr_max = 0.0f;
    // per tile (k & 3): its fixed-point step, its guessed thresholds (exponent audit)
    float q0 = 1.0f, q1 = 1.0f, q2 = 1.0f, q3 = 1.0f;
    float tl0 = 3.0e38f, tl1 = 3.0e38f, tl2 = 3.0e38f, tl3 = 3.0e38f, th0 = 0.0f, th1 = 0.0f, th2 = 0.0f, th3 = 0.0f;
    // guesses of tile j from a window sum known now, whose products with lo/L and hi/L are TLm and THm: chunk c is taken
    // to see that window sum + dr * (ofs + fc)
    auto prepare = [&](int j, float TLm, float THm, float dr, float ofs, float a_e, bool sane) {
        float q, invq;
        const bool stepok = pipe_step(a_e, q, invq);
        const float go = dr * (ofs + fc);
        const float TL = fmaf(go, loLf, TLm), TH = fmaf(go, hiLf, THm);
        const float cg = 0.5f * (TL + TH), rg = 0.5f * (TH - TL);
        if (actw) {
            ps.G[j & 3][lane] = make_float4(-cg, rg, invq, 0.0f);
            ps.J[j & 3][lane] = make_float4(go, TL, TH * 0x1p-19f, q);
        }
        const bool ok = __all_sync(FULL, !actw || (TL > 0.0f && rg > 0.0f)) && stepok && sane;
        if (lane == 0) ps.gok[j & 3] = ok ? 1 : 0;
        switch (j & 3) {
            case 0: q0 = q; tl0 = TL; th0 = TH; break;
            case 1: q1 = q; tl1 = TL; th1 = TH; break;
            case 2: q2 = q; tl2 = TL; th2 = TH; break;
            default: q3 = q; tl3 = TL; th3 = TH;
        }
    };
    float TLm = __double2float_rn(ssm * loL), THm = __double2float_rn(ssm * hiL);  // of ssm at the start of the tile being judged
    bool sane;  // the window sum at the start of the tile being judged is positive and in range
    {
        const float ssf = __double2float_rd(ssm);
        sane = ssf - hw0 * 2.0f > 0.0f && ssf < 1.0e30f && ssf > 1.0e-30f;
        prepare(0, TLm, THm, drift, 0.0f, a_est, sane);
        prepare(1, TLm, THm, drift, 1.0f, a_est, sane);
        prepare(2, TLm, THm, drift, 2.0f, a_est, sane);
    }
    if (lane == 0) {
        ps.vjudge[0] = ps.vjudge[1] = ps.vjudge2[0] = ps.vjudge2[1] = 0;
        ps.vst2[0] = ps.vst2[1] = ps.vst2b[0] = ps.vst2b[1] = 0;
    }
    // hysteresis carries
    int64_t lastL = cs.lastL, lrun_start = cs.lrun_start;
    int last_val = cs.last_val;
    const bool act = lane < C::NC;
    named_bar_sync<PIPE_BAR_RUN, (NW + 2) * 32>();  // barriers, guesses and flags are set up

    unsigned rph = 0u, vph = 0u;
    int k = 0, st = 0, redone = 0;
#pragma unroll 1
    for (; k < K; k++) {
        const int b = k & 1;
        const int64_t P0 = plan.tile0_pos + (int64_t)(t0 + k) * C::T;
        const float q = (k & 2) ? (b ? q3 : q2) : (b ? q1 : q0);
        bool st2 = false;
        int newL = -1, newS = -1, lv_new = 0;
        // the class maps of the tile as the workers left them in ps.bm[b]
        auto maps = [&]() {
            uint4 nl = make_uint4(FULL, FULL, FULL, FULL), hh = make_uint4(0u, 0u, 0u, 0u);
            if (act) {
                nl = *reinterpret_cast<const uint4 *>(&ps.bm[b][lane * 8]);
                hh = *reinterpret_cast<const uint4 *>(&ps.bm[b][lane * 8 + 4]);
            }
            const bool hasL = (nl.x & nl.y & nl.z & nl.w) != FULL, hasH = (hh.x | hh.y | hh.z | hh.w) != 0u;
            const int firstc = (int)((nl.x & 1u) + (hh.x & 1u)), lastc = (int)((nl.w >> 31) + (hh.w >> 31));
            const unsigned Lmask = __ballot_sync(FULL, hasL), Hmask = __ballot_sync(FULL, hasH);
            // hysteresis can matter only if a HIGH sample comes within max_len + 1 samples after a LOW sample
            st2 = false;
            if (Hmask) {
                const int nb = plan.nb;
                const int lo_c = max(lane - nb, 0);
                const unsigned win = (Lmask >> lo_c) & ((2u << (lane - lo_c)) - 1u);
                bool risk = hasH && win != 0u;
                if (hasH && lastL != NO_POS) {
                    const int64_t dist = P0 + (int64_t)lane * FAST_CH - lastL;  // first sample of the chunk to the carried LOW
                    if (dist <= (int64_t)mx + 1) risk = true;
                }
                st2 = __any_sync(FULL, risk);
            }
            // the carries the tile would leave: val of its last sample, last LOW sample and the start of its run
            newL = -1;
            newS = -1;
            if (Lmask) {
                int prevlast = __shfl_up_sync(FULL, lastc, 1);
                if (lane == 0) prevlast = last_val + 1;  // class code of the sample before the tile
                int candL = -1, candS = -1;
                if (hasL && (lastc == 0 || lane >= C::NC - plan.nb)) {
                    const unsigned NLw[4] = {nl.x, nl.y, nl.z, nl.w};
                    int bestL = -1, bestS = -1;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const unsigned lw = ~NLw[j];
                        const unsigned pw = j == 0 ? ((NLw[3] << 1) & ~1u) : NLw[j - 1];  // predecessor not LOW
                        const unsigned sw = lw & pw;
                        if (lw) bestL = max(bestL, ((31 - __clz(lw)) << 2) | j);
                        if (sw) bestS = max(bestS, ((31 - __clz(sw)) << 2) | j);
                    }
                    candL = lane * FAST_CH + bestL;
                    if (bestS >= 0) candS = lane * FAST_CH + bestS;
                    if (firstc == 0 && prevlast != 0) candS = max(candS, lane * FAST_CH);  // a LOW run starts at the chunk's first sample
                }
                newL = __reduce_max_sync(FULL, candL);
                newS = __reduce_max_sync(FULL, candS);
            }
            lv_new = __shfl_sync(FULL, lastc, C::NC - 1) - 1;
        };
        if (lane == 0) {  // for the tile's precise pass, should it need one
            ps.tlm[b] = TLm;
            ps.thm[b] = THm;
        }
        mbar_wait(&ps.rec_full[b], (rph >> b) & 1u);
        rph ^= 1u << b;
        maps();
        __syncwarp();
        if (lane == 0) {
            ps.vst2[b] = st2 ? 1 : 0;
            mbar_arrive(&ps.verdict[b]);
        }
        // ---------------------------------------------------------------- the mapper's say is out
        // the tile's drift and an upper bound of its sum of |x - prev|, from the workers' records; the window sum after the tile
        const PipeRec rc = ps.recs[b][actw ? lane : 0];
        int toti = __reduce_add_sync(FULL, actw ? rc.S : 0);
        float totA = __uint_as_float(__reduce_max_sync(FULL, actw ? __float_as_uint(rc.amax) : 0u)) * ((float)C::T * 1.0002f);
        float dr = (float)toti * q;
        float a_new = fmaxf(fmaxf(totA, 0.25f * a_est), TLm * 0x1p-16f);  // follows the traffic, decays slowly
        const float tl_k = (k & 2) ? (b ? tl3 : tl2) : (b ? tl1 : tl0), th_k = (k & 2) ? (b ? th3 : th2) : (b ? th1 : th0);
        double ssm2 = ssm + (double)toti * (double)q;  // exact product
        mbar_wait(&ps.verdict[b], (vph >> b) & 1u);  // the judge's say
        vph ^= 1u << b;
        int vj;
        {
            const volatile int *v = ps.vjudge;
            vj = v[b];
        }
        if (st2 || vj == PV_ABORT) break;
        bool reprime = false;
        if (vj == PV_REDO) {  // the precise pass: its maps, its sums; the coming two tiles are guessed from the window sum after it
            mbar_wait(&ps.rec_full[b], (rph >> b) & 1u);
            rph ^= 1u << b;
            const PipeRec r2 = ps.recs[b][actw ? lane : 0];
            toti = __reduce_add_sync(FULL, actw ? r2.S : 0);
            float incA2 = actw ? r2.amax : 0.0f;
#pragma unroll
            for (int o = NW / 2; o > 0; o >>= 1) incA2 += __shfl_xor_sync(FULL, incA2, o);
            totA = __shfl_sync(FULL, incA2, 0) * q * 1.0002f;
            dr = (float)toti * q;
            a_new = fmaxf(fmaxf(totA, 0.25f * a_est), TLm * 0x1p-16f);
            ssm2 = ssm + (double)toti * (double)q;
            const float ssfr = __double2float_rd(ssm2);
            const float TLr = __double2float_rn(ssm2 * loL), THr = __double2float_rn(ssm2 * hiL);
            const bool saner = ssfr * 0.99f > 0.0f && ssfr < 1.0e30f && ssfr > 1.0e-30f;
            prepare(k + 1, TLr, THr, dr, 0.0f, a_new, saner);
            prepare(k + 2, TLr, THr, dr, 1.0f, a_new, saner);
            prepare(k + 3, TLr, THr, dr, 2.0f, a_new, saner);
            maps();
            __syncwarp();
            if (lane == 0) {
                ps.vst2b[b] = st2 ? 1 : 0;
                mbar_arrive(&ps.verdict[b]);
            }
            mbar_wait(&ps.verdict[b], (vph >> b) & 1u);
            vph ^= 1u << b;
            const volatile int *v2 = ps.vjudge2;
            if (st2 || v2[b] != PV_ACCEPT) break;
            redone++;
            reprime = true;
        }
        // ---------------------------------------------------------------- tile k stands
        // admitted samples lie strictly between the guessed thresholds: one binade of slack either way (exponent audit)
        thr_min = fminf(thr_min, (reprime ? TLm : tl_k) * 0.5f);
        thr_max = fmaxf(thr_max, (reprime ? THm : th_k) * 2.0f);
        ssm = ssm2;
        const float ssf = __double2float_rd(ssm2);
        TLm = __double2float_rn(ssm2 * loL);
        THm = __double2float_rn(ssm2 * hiL);
        sane = ssf * 0.99f > 0.0f && ssf < 1.0e30f && ssf > 1.0e-30f;
        a_est = a_new;
        drift = dr;
        last_val = lv_new;
        if (newL >= 0) {
            lastL = P0 + newL;
            if (newS >= 0) lrun_start = P0 + newS;
        }
        // the guesses of tile k + 3, from the window sum after this tile: this tile's drift, as much again for each of the two
        // tiles in between (published by the verdict on tile k + 1, before any worker starts tile k + 3)
        if (!reprime) prepare(k + 3, TLm, THm, dr, 2.0f, a_new, sane);
        if (reprime && k + 1 < K) issue(k + 1, st + 1 == S ? 0 : st + 1);  // its samples again: its stage holds the log of the tile taken back
        // its stage (the undo log by now) is free for tile k + S
        if (issued < K) {
            issue(issued, st);
            issued++;
        }
        if (++st == S) st = 0;
    }
    // ---- hand the state back: the window sum after the last proven tile (the judge adds the interval's half width)
    for (int s = 0; s < S; s++)  // no copy in flight
        if (xused & (1u << s)) mbar_wait(&ps.x_full[s], (xlast >> s) & 1u);
    thr_min = redux_min(thr_min);
    thr_max = -redux_min(-thr_max);
    if (lane == 0) {
        ps.ssm_out = ssm;
        if (k > 0) {
            uni.tot_prev = drift;
            uni.a_est = a_est;
            uni.thr_min = fminf(uni.thr_min, thr_min);
            uni.thr_max = fmaxf(uni.thr_max, thr_max);
            uni.stats[FS_FAST] += (unsigned)k;
            uni.stats[FS_PIPE_T] += (unsigned)k;
            uni.stats[FS_PIPE_RD] += (unsigned)redone;
        }
        cs.lastL = lastL;
        cs.lrun_start = lrun_start;
        cs.last_val = last_val;
    }
}

// IQ input has no pipelined mode (a tile of complex samples does not fit beside the ring twice): the extra warps only leave
template <int NW>
__device__ __forceinline__ void pipe_aux_idle_impl(volatile int *cmd) {
    for (;;) {
        named_bar_sync<PIPE_BAR_PARK, (NW + 2) * 32>();
        if (*cmd == PIPE_CMD_QUIT) return;
    }
}
template <int NW, class PS>
__device__ __forceinline__ void pipe_aux_idle(PS &ps) {
    pipe_aux_idle_impl<NW>(&ps.cmd);
}

// The two extra warps of a CTA: parked until the segment's workers enter the pipelined mode (or finish the segment).
template <int NW, int R, int S, int ITEM>
__device__ __forceinline__ void pipe_aux_main(PipeShared<NW, R, S> &ps, FastUni &uni, const FastPlan &plan, const SlicerParams &p, SegCarry &cs,
                                              float *ring, const int warp, const int lane, const int allow_redo) {
    for (;;) {
        named_bar_sync<PIPE_BAR_PARK, (NW + 2) * 32>();
        if (*(volatile int *)&ps.cmd == PIPE_CMD_QUIT) return;
        char *stage0 = reinterpret_cast<char *>(ring) + (((size_t)p.L * 4 + 15) / 16) * 16;  // the stages follow the ring
        if (warp == NW) pipe_judge<NW, R, S>(ps, uni, plan, lane, allow_redo);
        else pipe_mapper<NW, R, S, ITEM>(ps, uni, cs, plan, p.loL, p.hiL, stage0, p.mx, lane);
        named_bar_sync<PIPE_BAR_RUN, (NW + 2) * 32>();  // the run is over, the state handed back
    }
}

}  // namespace nfc
