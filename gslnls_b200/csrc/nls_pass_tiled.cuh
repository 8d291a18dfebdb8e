// nls_pass_tiled.cuh -- K1b: the fused pass for p > 8, J^T J by FP64 tensor-core SYRK.
//
// Included by nls_pass_kernel.cuh when NLS_TILED=1 (same entry point name, same packet).  With
// p(p+1)/2 in the hundreds the normal-equation accumulators no longer fit a thread's registers,
// so the work moves in slabs of 32 observations through shared-memory tiles, between two kinds of
// warps of the same CTA (warp specialisation; hand-off through mbarriers, no CTA-wide barrier in
// the streaming loop):
//
//   producer warps (phase A)  lane i evaluates residual and Jacobian row of observation i (the model
//            is inlined) and parks the sqrt(w)-scaled row in the warp's tile, transposed:
//            tile[column][observation].  They never hold accumulators, so the row code has the
//            registers it needs.
//   consumer warps (phase B)  one per SM sub-partition, serving NT_NPROD producers in a fixed
//            round-robin; they hold the T accumulator fragments for the whole launch and never
//            evaluate the model, so nothing spills.  A consumer contracts a tile with itself: 8
//            k-steps of 4 observations, each feeding
//            one mma.sync.m8n8k4.f64 (DMMA) per 8x8 block pair of the lower triangle.  A- and
//            B-fragments of a parameter block are the same register, so a k-step costs
//            ceil(p/8) shared loads for T = PB(PB+1)/2 DMMAs; the T accumulator fragments stay
//            in registers for the whole launch (42 doubles per lane at p = 48).
//            J^T r (or J^T fvv, J^T (J d)) reuses the fragment registers: one FMA per block and
//            k-step into per-lane partial sums, folded across the four fragment columns at the end.
//
// Replaces for large p what nls_pass does for small p: gsl_df_large's dsyrk/dgemv
// (src/nls_large.c:629,633) without ever forming the n x p matrix J (src/nls_large.c:167).
// Nothing of size n is written.  On B200 DMMA issues to the same FP64 units as DFMA (measured:
// 37.1 TFLOP/s either way, no overlap), so the roofline of this kernel is the FP64 pipe; what the
// tensor-core form buys is one instruction and two operand registers per 256 FMAs.
//
// Both kinds of warp issue to the same FP64 pipe; the consumers always have independent DMMAs ready,
// which fills the latency bubbles of the producers' dependent chains.
//
// Tunables: NLS_UNROLL = NT_NPROD producer warps per consumer warp, NLS_BLOCK = 32 * NT_NCONS *
// (1 + NT_NPROD) threads per CTA, NLS_MINB CTAs per SM.

#define NT_PB ((NLS_P + 7) / 8)          /* 8-wide parameter blocks                         */
#define NT_COLS (NT_PB * 8)              /* J columns incl. zero padding                    */
#define NT_LDT 32                        /* tile column pitch in doubles: the 32 observations of a slab, no padding.
                                            Element (column j, observation i) sits at NT_AT(j, i): the observation
                                            index is XOR-swizzled with the column's low two bits, so both the column
                                            stores (a permutation of 32 consecutive doubles) and the 4x8 fragment
                                            loads (columns fr, observations 4 ks + fc) hit 16 distinct 8-byte banks
                                            per half-warp.  (A padded pitch of 36 does the same with 12 % more
                                            shared memory -- one tile buffer per CTA at p = 48.) */
#define NT_AT(j, i) ((j) * NT_LDT + ((i) ^ (((j) & 3) << 2)))
#define NT_TILES (NT_PB * (NT_PB + 1) / 2)
#define NT_GC NT_PB                      /* column-dot partial sums per lane (one per block)  */

#ifndef NT_STAGE_REGS
#define NT_STAGE_REGS 0 /* 1: consumers release a tile after half of its DMMAs (second half staged in registers).
                           Measured at p = 48, n = 1e7: no gain (1354.6 vs 1353.3 us per pass); staging the whole
                           tile spills at 128 registers per thread and costs 45 % (1968.8 us).  The coupling that is
                           left -- consumers wait 19 % of their time for tiles, producers 20 % for an empty buffer,
                           one buffer per producer, 11.5 k cycles to produce a slab against 0.8 k of FP64 issue --
                           needs more tile buffers than 227 KB of shared memory hold at this tile format. */
#endif
#define NT_NPROD NLS_UNROLL
#define NT_NCONS (NLS_NW / (1 + NT_NPROD))
#define NT_NP (NT_NCONS * NT_NPROD)      /* producer warps (= tile buffers) per CTA           */
#define NT_TILE_DOUBLES (NT_COLS * NT_LDT + 32)
// Tile buffers are decoupled from the producers: every consumer owns a ring of R_c buffers that its NT_NPROD
// producers fill in the consumer's own consumption order (slab number Lc = r NT_NPROD + j of round r, producer
// j lives in buffer base_c + Lc mod R_c).  With one buffer per producer (R_c = NT_NPROD, the round-1 layout) a
// producer cannot start its next slab before its previous one has been consumed, and both roles measured ~20 %
// of their time waiting for each other; the NT_NBUF buffers that 227 KB of shared memory hold (NT_NP + 3 at
// p = 48) are dealt out 4,4,4,3.  The rings are per consumer because mbarrier waits only see the phase PARITY:
// a waiter must never be two phases away from its barrier.  Inside one consumer's ring that holds by
// construction -- the consumer takes its slabs in order, so a producer that has filled its previous slab
// Lc - NT_NPROD knows every slab up to Lc - NT_NPROD - R_c has been consumed, and Lc - 2 R_c is among them when
// R_c >= NT_NPROD -- whereas a ring shared by independent consumers can alias (a fast consumer's producer would
// read "free" off a buffer a slow consumer is two uses behind on).  The order in which a consumer contracts
// slabs -- hence every sum -- does not depend on the ring sizes.  The host passes NT_NBUF (model.cpp).
#ifndef NT_NBUF
#define NT_NBUF NT_NP
#endif
#if NT_NBUF / NT_NCONS < NT_NPROD
#error "nls_pass_tiled: every consumer needs at least one tile buffer per producer"
#endif
// ring position of round r of producer q: buffer and how many times that buffer has been used before
static __device__ __forceinline__ void nt_ring_pos(int q, unsigned r, unsigned &buf, unsigned &use)
{
    const int c = q / NT_NPROD, j = q - c * NT_NPROD;
    const int RB = NT_NBUF / NT_NCONS, RX = NT_NBUF % NT_NCONS;
    const unsigned R = (unsigned)(RB + (c < RX ? 1 : 0)), base = (unsigned)(c * RB + (c < RX ? c : RX));
    const unsigned Lc = r * NT_NPROD + (unsigned)j;
    buf = base + Lc % R;
    use = Lc / R;
}

#if NT_TILES > 28
#error "nls_pass_tiled: p > 56 needs the tile set split across warps (not built yet)"
#endif

// ---- mbarrier hand-off (shared::cta): 32 arrivals per phase, parity waits ----
static __device__ __forceinline__ unsigned nt_saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ void nt_bar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nt_saddr(bar)), "r"(count) : "memory");
}
static __device__ __forceinline__ void nt_bar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(nt_saddr(bar)) : "memory");
}
static __device__ __forceinline__ void nt_bar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "NT_WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra NT_DONE_%=;\n"
                 "bra NT_WAIT_%=;\n"
                 "NT_DONE_%=:\n"
                 "}" ::"r"(nt_saddr(bar)), "r"(parity) : "memory");
}

static __device__ __forceinline__ void nt_dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

#if GSLNLS_JAC_MODE == 0
#define NT_NC_FJ GSLNLS_NC_FJ
#else
#define NT_NC_FJ 0
#endif
#if GSLNLS_FVV_MODE == 1
#define NT_NC_FVV GSLNLS_NC_FVV
#else
#define NT_NC_FVV 0
#endif
// the slow paths (finite-difference Jacobian / fvv) keep parameters and steps in registers
#define NT_NEED_T (GSLNLS_JAC_MODE != 0 || GSLNLS_FVV_MODE == 2)

// parameters, velocity and the launch invariants of the generated model live in shared memory
struct NtShared {
    const double *th, *vv, *c_fj, *c_fvv;
};

// receives Jacobian entries from the generated row function as they are produced: scaled by
// sqrt(w) (0 for the padding lanes of the last slab) straight into this lane's tile column; the
// JVP mode also needs u = row . d, accumulated on the way
struct NtRowSink {
    double *tile;
    int lane;
    const double *vv;
    double sw, u;
    __device__ __forceinline__ void operator()(int j, double v)
    {
        const double s = v * sw;
        tile[NT_AT(j, lane)] = s;
        u = fma(s, vv[j], u);
    }
};
struct NtRowSinkNoDot {
    double *tile;
    int lane;
    double sw;
    __device__ __forceinline__ void operator()(int j, double v) { tile[NT_AT(j, lane)] = v * sw; }
};

// ---------------------------------------------------------------- producer: phase A
template <int MODE>
static __device__ __forceinline__ void nt_produce(const NlsPassParams &prm, const NlsThread &T, const NtShared &S,
                                                  int q, double *tiles, unsigned long long *full_base,
                                                  unsigned long long *empty_base, double &ss, double &nbad,
                                                  long long &nt_wait_cycles)
{
    const int lane = threadIdx.x & 31;
    const long long n = prm.n;
    const long long nslab = (n + 31) >> 5;
    const long long wstride = (long long)gridDim.x * NT_NP;
    // software prefetch: the predictor / response / weight of the next slab are requested before this
    // slab's arithmetic starts
    long long slab = (long long)blockIdx.x * NT_NP + q;
    double nx[NLS_NV], ny = 0.0, nw = 1.0;
    {
        const long long i0 = (slab << 5) + lane;
        const long long ii0 = (slab < nslab && i0 < n) ? i0 : 0;
#pragma unroll
        for (int k = 0; k < GSLNLS_NVAR; ++k)
            nx[k] = nls_ld1(prm.vars[k] + ii0);
        ny = nls_ld1(prm.y + ii0);
#if NLS_HAS_W
        nw = nls_ld1(prm.w + ii0);
#endif
    }
    int nbad_i = 0;
    for (unsigned r = 0; slab < nslab; slab += wstride, ++r) {
        unsigned buf, use;
        nt_ring_pos(q, r, buf, use);
        double *tile = tiles + (size_t)buf * NT_TILE_DOUBLES;
        double *rv = tile + NT_COLS * NT_LDT;
        unsigned long long *full = full_base + buf, *empty = empty_base + buf;
        const long long i = (slab << 5) + lane;
        const bool valid = i < n;
        double xa[NLS_NV];
#pragma unroll
        for (int k = 0; k < GSLNLS_NVAR; ++k)
            xa[k] = nx[k];
        const double y = ny;
#if NLS_HAS_W
        const double sw = sqrt(nw); // sqrt_wts_i = sqrt(w_i), src/fdf.c:60-64
#else
        const double sw = 1.0;
#endif
        {
            const long long in = ((slab + wstride) << 5) + lane;
            const long long iin = (slab + wstride < nslab && in < n) ? in : 0;
#pragma unroll
            for (int k = 0; k < GSLNLS_NVAR; ++k)
                nx[k] = nls_ld1(prm.vars[k] + iin);
            ny = nls_ld1(prm.y + iin);
#if NLS_HAS_W
            nw = nls_ld1(prm.w + iin);
#endif
        }
        {
            const long long w0 = prm.trace ? clock64() : 0;
            nt_bar_wait(empty, (use & 1u) ^ 1u); // the consumer is done with this buffer's previous contents
            if (prm.trace)
                nt_wait_cycles += clock64() - w0;
        }
        // sqrt(w) scaling folded into the tile store; invalid (padding) lanes store zeros
#if NLS_W_GSL
        const double swv = valid ? 1.0 : 0.0; // reference-compatible weights: the rows of J stay unweighted
#else
        const double swv = valid ? sw : 0.0;
#endif
        double f, u = 0.0;
#if defined(NT_DEBUG_SKIP_A)
        f = xa[0]; // timing experiment: no model evaluation, tile keeps its initial contents
#elif GSLNLS_JAC_MODE == 0
        if (MODE == NLS_MODE_JVP || (MODE == NLS_MODE_FVV && GSLNLS_FVV_MODE == 2)) {
            NtRowSink sink{tile, lane, S.vv, swv, 0.0};
            nls_model_fj_c(S.th, S.c_fj, xa, f, sink);
            u = sink.u;
        } else {
            NtRowSinkNoDot sink{tile, lane, swv};
            nls_model_fj_c(S.th, S.c_fj, xa, f, sink);
        }
#else
        double J[NLS_P];
        nls_fj(T, xa, f, J);
#pragma unroll
        for (int j = 0; j < NLS_P; ++j) {
            tile[NT_AT(j, lane)] = J[j] * swv;
            u = fma(J[j] * swv, S.vv[j], u);
        }
#endif
        double rr;
        if (MODE == NLS_MODE_FJ) {
            const bool bad = !nls_finite(f); // -> residual +Inf, src/nls_large.c:464-465 (selects, no branch)
            rr = (bad ? NLS_INF : f - y) * sw;
            nbad_i += (bad && valid) ? 1 : 0;
        } else if (MODE == NLS_MODE_FVV) {
#if GSLNLS_FVV_MODE == 1
            rr = nls_model_fvv_c(S.th, S.vv, S.c_fvv, xa) * sw;
#elif GSLNLS_FVV_MODE == 2
            {
                // fvv = (2/h) ((f(x + h v) - f(x)) / h - J v), src/fdfvv.c:47-74  (u = sqrt(w) J v here)
                double tp[NLS_P];
#pragma unroll
                for (int k = 0; k < NLS_P; ++k)
                    tp[k] = T.th[k] + T.h_fvv * T.vv[k];
                const double fp = nls_model_f(tp, xa);
                const double hinv = 1.0 / T.h_fvv;
#if NLS_W_GSL
                rr = (2.0 * hinv) * ((fp - f) * hinv - u) * sw; // u = J v, unweighted rows
#else
                rr = (2.0 * hinv) * (((fp - f) * hinv) * sw - u);
#endif
            }
#else
            rr = 0.0;
#endif
        } else { // JVP: u = (sqrt(w) J) d
            rr = u;
        }
        if (!valid)
            rr = 0.0;
        rv[lane] = rr;
        ss = fma(rr, rr, ss);
        nt_bar_arrive(full); // 32 arrivals (release): the tile and rv are complete
    }
    nbad = (double)nbad_i;
}

// ---------------------------------------------------------------- consumer: phase B
template <int MODE>
static __device__ __forceinline__ void nt_consume(const NlsPassParams &prm, int c, double *tiles,
                                                  unsigned long long *full, unsigned long long *empty,
                                                  double (&C)[NT_TILES][2], double (&gacc)[NT_GC],
                                                  long long &nt_wait_cycles)
{
    const int lane = threadIdx.x & 31;
    const long long nslab = (prm.n + 31) >> 5;
    const long long wstride = (long long)gridDim.x * NT_NP;
    const int fr = lane >> 2, fc = lane & 3; // fragment coordinates of this lane
    // NT_AT(8 b + fr, 4 ks + fc) = 8 b NT_LDT + frow + ((4 ks) ^ fx): fc sits below the swizzled bits
    const int frow = fr * NT_LDT + fc, fx = (fr & 3) << 2;
    // producers q = c NT_NPROD + j, j = 0..NT_NPROD-1, each in its own slab sequence; the consumer
    // takes their tiles in the fixed order (round 0: j = 0, 1, ..; round 1: ..) -- deterministic sums
    for (unsigned r = 0;; ++r) {
        bool any = false;
#pragma unroll 1
        for (int j = 0; j < NT_NPROD; ++j) {
            const int q = c * NT_NPROD + j;
            const long long slab = (long long)blockIdx.x * NT_NP + q + (long long)r * wstride;
            if (slab >= nslab)
                continue;
            any = true;
            unsigned buf, use;
            nt_ring_pos(q, r, buf, use);
            const double *tile = tiles + (size_t)buf * NT_TILE_DOUBLES;
            const double *rv = tile + NT_COLS * NT_LDT;
            {
                const long long w0 = prm.trace ? clock64() : 0;
                nt_bar_wait(full + buf, use & 1u);
                if (prm.trace)
                    nt_wait_cycles += clock64() - w0;
            }
            // k-step ks covers observations 4 ks .. 4 ks + 3; this lane's fragment element of block b is
            // J[observation 4 ks + fc][column 8 b + fr].  The same registers feed the column dot with r
            // (J^T r, or J^T fvv / J^T (J d)): a partial sum per lane, folded over fc once per launch.
#if NT_STAGE_REGS
            // The tile is handed back to its producer after HALF of the slab's DMMAs, not after all of them:
            // the second half of the tile (k-steps 4..7) moves into registers first.  (Staging the whole tile
            // up front needs 56 more registers than the 128 a 512-thread CTA leaves per thread: it spills and
            // the consumer becomes the bottleneck, 1.97 ms per pass against 1.35 ms -- measured.)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                double fragh[4][NT_PB], rkh[4];
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const int ks = half * 4 + k4;
#pragma unroll
                    for (int b = 0; b < NT_PB; ++b)
                        fragh[k4][b] = tile[b * 8 * NT_LDT + frow + ((ks * 4) ^ fx)];
                    rkh[k4] = rv[ks * 4 + fc];
                }
                if (half == 1)
                    nt_bar_arrive(empty + buf); // every load of this tile has been issued before (release)
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
#if !defined(NT_DEBUG_SKIP_B)
                    if (MODE == NLS_MODE_FJ) {
                        int t = 0;
#pragma unroll
                        for (int bi = 0; bi < NT_PB; ++bi)
#pragma unroll
                            for (int bj = 0; bj <= bi; ++bj, ++t)
                                nt_dmma(C[t][0], C[t][1], fragh[k4][bi], fragh[k4][bj]);
                    }
#endif
#pragma unroll
                    for (int b = 0; b < NT_PB; ++b)
                        gacc[b] = fma(fragh[k4][b], rkh[k4], gacc[b]);
                }
            }
#else
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                double frag[NT_PB];
#pragma unroll
                for (int b = 0; b < NT_PB; ++b)
                    frag[b] = tile[b * 8 * NT_LDT + frow + ((ks * 4) ^ fx)];
                const double rk = rv[ks * 4 + fc];
                if (ks == 7)
                    nt_bar_arrive(empty + buf); // every load of this tile has been issued before (release)
#if !defined(NT_DEBUG_SKIP_B)
                if (MODE == NLS_MODE_FJ) {
                    int t = 0;
#pragma unroll
                    for (int bi = 0; bi < NT_PB; ++bi)
#pragma unroll
                        for (int bj = 0; bj <= bi; ++bj, ++t)
                            nt_dmma(C[t][0], C[t][1], frag[bi], frag[bj]);
                }
#endif
#pragma unroll
                for (int b = 0; b < NT_PB; ++b)
                    gacc[b] = fma(frag[b], rk, gacc[b]);
            }
#endif
        }
        if (!any)
            break;
    }
}

extern "C" __global__ void __launch_bounds__(NLS_BLOCK, NLS_MINB) nls_pass(const NlsPassParams prm)
{
    extern __shared__ double nt_smem[];
    const int cand = blockIdx.y;
    const double *req = prm.req + (size_t)cand * prm.req_stride;
    const int mode = nls_begin(prm, req);
    if (mode == NLS_MODE_IDLE)
        return;
    nls_exp_init();

    NlsThread T;
#if NT_NEED_T
    nls_load_request(prm, req, T);
#endif

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // shared memory: NT_NP tiles [NT_COLS][NT_LDT] + rv[32]; the CTA packet [NLS_PK]; theta, v and the
    // model's launch invariants; the full / empty barriers of every tile
    double *tiles = nt_smem;
    double *pk = nt_smem + (size_t)NT_NBUF * NT_TILE_DOUBLES;
    double *s_th = pk + NLS_PK, *s_vv = s_th + NLS_P, *s_cfj = s_vv + NLS_P, *s_cfvv = s_cfj + NT_NC_FJ;
    unsigned long long *full = (unsigned long long *)(s_cfvv + NT_NC_FVV), *empty = full + NT_NBUF;
    for (int e = threadIdx.x; e < NT_NBUF * NT_TILE_DOUBLES; e += NLS_BLOCK)
        tiles[e] = 0.0; // padding columns stay zero for the whole launch
    for (int e = threadIdx.x; e < NLS_PK; e += NLS_BLOCK)
        pk[e] = 0.0;
    for (int j = threadIdx.x; j < NLS_P; j += NLS_BLOCK) {
        s_th[j] = __ldcg(req + 1 + j);
        s_vv[j] = __ldcg(req + 1 + NLS_P + j);
    }
    if (threadIdx.x < NT_NBUF) {
        nt_bar_init(full + threadIdx.x, 32);
        nt_bar_init(empty + threadIdx.x, 32);
    }
    __syncthreads();
#if GSLNLS_JAC_MODE == 0
    if (threadIdx.x == 0)
        nls_model_prep_fj(s_th, s_cfj);
#endif
#if GSLNLS_FVV_MODE == 1
    if (threadIdx.x == NLS_BLOCK - 1 && mode == NLS_MODE_FVV)
        nls_model_prep_fvv(s_th, s_vv, s_cfvv);
#endif
    __syncthreads();
    NtShared S;
    S.th = s_th; S.vv = s_vv; S.c_fj = s_cfj; S.c_fvv = s_cfvv;

    double C[NT_TILES][2];
#pragma unroll
    for (int t = 0; t < NT_TILES; ++t)
        C[t][0] = C[t][1] = 0.0;
    double gacc[NT_GC];
#pragma unroll
    for (int g = 0; g < NT_GC; ++g)
        gacc[g] = 0.0;
    double ss = 0.0, nbad = 0.0;

    const bool consumer = warp < NT_NCONS;
    long long nt_wait_cycles = 0; // developer hook: cycles this warp spent waiting for the other role
    const long long nt_t0 = prm.trace ? clock64() : 0;
    if (consumer) {
        if (mode == NLS_MODE_FJ)
            nt_consume<NLS_MODE_FJ>(prm, warp, tiles, full, empty, C, gacc, nt_wait_cycles);
        else
            nt_consume<NLS_MODE_FVV>(prm, warp, tiles, full, empty, C, gacc, nt_wait_cycles); // FVV and JVP: column dots only
    } else {
        const int q = warp - NT_NCONS;
        if (mode == NLS_MODE_FJ)
            nt_produce<NLS_MODE_FJ>(prm, T, S, q, tiles, full, empty, ss, nbad, nt_wait_cycles);
        else if (mode == NLS_MODE_FVV)
            nt_produce<NLS_MODE_FVV>(prm, T, S, q, tiles, full, empty, ss, nbad, nt_wait_cycles);
        else
            nt_produce<NLS_MODE_JVP>(prm, T, S, q, tiles, full, empty, ss, nbad, nt_wait_cycles);
    }
    if (prm.trace && lane == 0 && warp < 16 && blockIdx.y == 0) {
        // [cta][0..15]: wait cycles of warp w; [cta][16 + w]: cycles warp w spent in its streaming loop
        prm.trace[(size_t)blockIdx.x * 32 + warp] = (unsigned long long)nt_wait_cycles;
        prm.trace[(size_t)blockIdx.x * 32 + 16 + warp] = (unsigned long long)(clock64() - nt_t0);
    }

    // ---- CTA reduction: warps add their sums into the shared packet in warp order ----
    ss += __shfl_down_sync(0xffffffffu, ss, 16);
    ss += __shfl_down_sync(0xffffffffu, ss, 8);
    ss += __shfl_down_sync(0xffffffffu, ss, 4);
    ss += __shfl_down_sync(0xffffffffu, ss, 2);
    ss += __shfl_down_sync(0xffffffffu, ss, 1);
    nbad += __shfl_down_sync(0xffffffffu, nbad, 16);
    nbad += __shfl_down_sync(0xffffffffu, nbad, 8);
    nbad += __shfl_down_sync(0xffffffffu, nbad, 4);
    nbad += __shfl_down_sync(0xffffffffu, nbad, 2);
    nbad += __shfl_down_sync(0xffffffffu, nbad, 1);
    const int fr = lane >> 2, fc2 = (lane & 3) * 2;
    // column dots: fold the four fc-lanes of each fragment row (fixed xor tree); lanes with fc == 0 own
    // column 8 b + fr afterwards
#pragma unroll
    for (int b = 0; b < NT_PB; ++b) {
        gacc[b] += __shfl_xor_sync(0xffffffffu, gacc[b], 1);
        gacc[b] += __shfl_xor_sync(0xffffffffu, gacc[b], 2);
    }
    const int goff = mode == NLS_MODE_FJ ? NLS_NPK : 0; // FVV / JVP packet: [J^T h (p) | h^T h]
    __syncthreads();
    for (int w = 0; w < NLS_NW; ++w) {
        if (warp == w) {
            if (consumer) {
                if (mode == NLS_MODE_FJ) {
                    // packet: [J^T J lower packed row-major | J^T f | f^T f | #non-finite]
                    int t = 0;
#pragma unroll
                    for (int bi = 0; bi < NT_PB; ++bi)
#pragma unroll
                        for (int bj = 0; bj <= bi; ++bj, ++t) {
                            const int row = bi * 8 + fr;
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int col = bj * 8 + fc2 + h;
                                if (row < NLS_P && col <= row)
                                    pk[row * (row + 1) / 2 + col] += C[t][h];
                            }
                        }
                }
#pragma unroll
                for (int b = 0; b < NT_PB; ++b) {
                    const int c = b * 8 + fr;
                    if ((lane & 3) == 0 && c < NLS_P)
                        pk[goff + c] += gacc[b];
                }
            } else if (lane == 0) {
                pk[goff + NLS_P] += ss;
                if (mode == NLS_MODE_FJ)
                    pk[goff + NLS_P + 1] += nbad;
            }
        }
        __syncthreads();
    }
    double *part = prm.partials + ((size_t)cand * gridDim.x + blockIdx.x) * prm.pk_stride;
    for (int e = threadIdx.x; e < NLS_PK; e += NLS_BLOCK)
        part[e] = pk[e];
    nls_grid_finish(prm, cand);
}
