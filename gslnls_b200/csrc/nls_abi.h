// nls_abi.h -- plain-C structs shared by the host library (nvcc/g++) and the NVRTC-compiled
// kernels.  Keep it free of includes so that NVRTC can consume it as an in-memory header.
#pragma once

#define NLS_MAX_VARS 16
#define NLS_MAX_RANKS 8

// Device-resident synchronisation channel between the pass kernel (K1), the resident trust-region
// server (trs_server, trs_kernel.cu) and -- over NVLink peer memory -- the other GPUs' pass kernels.
// Byte offsets into one cudaMalloc'd block (each word on its own 128-byte line); the block of every
// rank is mapped into every other rank's address space (cudaIpc), so a pass kernel deposits its
// packet directly into each peer's mailbox and no collective call sits between pass and step.
#define NLS_CH_REQ_SEQ 0      /* u64: requests published so far (server -> K1)                  */
#define NLS_CH_PASS_CTR 128   /* u64: passes completed so far (last CTA of K1 -> next K1)       */
#define NLS_CH_FIT_SEQ0 256   /* u64: sequence number of the first pass of the current fit      */
#define NLS_CH_ABORT 384      /* u64: host -> server, leave the fit early                       */
#define NLS_CH_TIMER 512      /* u64[5]: globaltimer ns, first CTA past the wait / last CTA out  */
#define NLS_CH_PASS_SEEN 640  /* u64: sequence number of the latest pass kernel that has STARTED running (K1 ->
                                 server): the start-of-fit handshake that proves pass and server kernels execute
                                 concurrently (they do not under ncu, compute-sanitizer, CUDA_LAUNCH_BLOCKING) */
#define NLS_CH_CTA_COUNT 768  /* u64: partial packets deposited by the CTAs of the persistent pass kernel since the
                                 fit began (K1 -> server): pass j of the fit is complete at j x gridDim.x */
#define NLS_CH_GONE_BUMP (1ull << 40) /* added to REQ_SEQ by a server that gives up: every queued pass falls through idle */
#define NLS_CH_FLAGS 1024     /* u64 x NLS_MAX_RANKS, 128 B apart: latest sequence deposited by rank r */
#define NLS_CH_DATA 2048      /* double [2 parities][NLS_MAX_RANKS][NLS_CH_MAXPK]               */
#define NLS_CH_MAXPK 1280
#define NLS_CH_BYTES (NLS_CH_DATA + 2 * NLS_MAX_RANKS * NLS_CH_MAXPK * 8)

// Parameters of one fused pass (K1).  Passed by value to the kernel.
struct NlsPassParams {
    const double *vars[NLS_MAX_VARS]; // nvar predictor columns, n doubles each (device)
    const double *y;                  // responses
    const double *w;                  // raw weights or nullptr
    long long n;                      // local observations
    const double *req;                // [ncand][req_stride]: mode, theta[p], v[p]
    double *partials;                 // [ncand][gridDim.x][pk_stride] per-CTA partial packets
    double *packet;                   // [ncand][pk_stride] reduced packet
    unsigned int *ticket;             // [ncand] arrival counters (zero between launches)
    double h_df, h_fvv;               // finite-difference steps (control_dbl[3], [4])
    int req_stride, pk_stride;
    int force_mode;                   // >0: ignore req[0] and run this mode (test hooks)
    int nranks;                       // server mode: ranks depositing into each mailbox (1 = this GPU only)
    // server mode (channel != nullptr): wait for the request, deposit the packet in every rank's
    // mailbox; launch-ordered mode (nullptr): the packet goes to `packet` and the host sequences
    char *channel;                    // this GPU's channel block
    char *peer_channel[NLS_MAX_RANKS];// every rank's channel block as mapped on this GPU ([rank] == channel)
    int rank;
    int pad_;
    int *prof_flag;                   // benchmark hook: set to 1 by a launch that really streamed (not idle)
    unsigned long long watchdog_ns;   // server mode: longest in-kernel wait for a request (GSLNLS_WATCHDOG_S)
    int server_reduce;                // persistent kernel: 1 = the CTAs only deposit partials and count themselves in
                                      // on NLS_CH_CTA_COUNT; the resident server sums them (no last-CTA stage)
    int pad3_;
    int max_passes;                   // persistent kernel: leave after this many passes even if the fit goes on
    int l2_ahead;                     // persistent kernel: tiles per CTA requested into L2 ahead of the ring
    volatile int *host_passes;        // persistent kernel: passes executed, written at exit (mapped host memory) or nullptr
    unsigned long long *trace;        // developer hook: [gridDim.x][8] globaltimer stamps of each CTA's phases, or nullptr
    long long keep_rows;              // rows [0, keep_rows) of every column are loaded with an L2 evict_last policy,
                                      // the rest evict_first: the head of the shard stays L2-resident from pass to pass
    // two-level grid reduction (single-candidate launches): CTAs in groups of NLS_RED_GROUP, the last
    // arriver of a group sums the group's partials, the last group sums the group sums
    double *group_partials;           // [ceil(gridDim.x / NLS_RED_GROUP)][pk_stride] or nullptr (flat reduction)
    unsigned int *group_ticket;       // [ceil(gridDim.x / NLS_RED_GROUP)] arrival counters (zero between launches)
};
#define NLS_RED_GROUP 32

struct NlsMaterialiseParams {
    const double *vars[NLS_MAX_VARS];
    const double *y;
    const double *w;
    long long n;
    const double *theta;  // p doubles (device)
    double *resid;        // n or nullptr
    double *grad;         // n*p column-major or nullptr
    double h_df;
};

// IRLS support kernels (robust losses, src/nls_irls.c): unweighted residuals r_i = fn(theta)_i - y_i are
// recomputed from the resident columns in every pass -- nothing of size n besides the weights column is written
struct NlsIrlsParams {
    const double *vars[NLS_MAX_VARS];
    const double *y;
    long long n;
    const double *theta;          // p doubles (device)
    // radix select of the k-th smallest |r| (bit pattern of a non-negative double is monotone)
    unsigned long long prefix;    // bits above `shift + 8` already fixed
    int shift;                    // current digit: bits [shift, shift + 8)
    int pad_;
    unsigned long long *hist;     // [256] counts of the current digit among the keys matching the prefix
    unsigned long long pivot;     // count / min pass: key of the order statistic found
    unsigned long long *cnt_min;  // [2]: #{key <= pivot}, min{key > pivot}
    // weights pass
    double sigma;                 // scale: 1.4826 median |r|
    int loss;                     // 1 huber 2 barron 3 bisquare 4 welsh 5 optimal 6 hampel 7 ggw 8 lqq
    int pad2_;
    double cc[3];                 // tuning constants of the loss (R/nls_rho.R)
    double *wout;                 // n: max(psi(r/sigma) / (r/sigma), eps), before normalisation
    double *psi, *psip;           // n each or nullptr
    double *partial;              // [gridDim.x] per-CTA sums of wout (summed in CTA order on the host)
    // scale pass: wout[i] *= scale * (userw ? userw[i] : 1)
    double scale;
    const double *userw;
    double h_df;
};

// Sparse-row problems (src/nls_large.c:528-648: the model returns J as a dgT/dgC/dgRMatrix).  A row of such a
// model touches a handful of the P parameters; the compiled row function sees them as its local parameters
// th[0..k), gathered per TERM from the global vector: th[s] = theta[slot_base[s] + (slot_index[s] ?
// slot_index[s][t] : 0)].  One launch evaluates every term of a block and stores its value and its k partial
// derivatives -- the nonzeros of J, slot-major so that every store is coalesced.
#define NLS_SP_MAXSLOT 16
struct NlsSparseEvalParams {
    const double *vars[NLS_MAX_VARS];        // data columns of the block, nterms doubles each
    const int *slot_index[NLS_SP_MAXSLOT];   // index column of local parameter s, or nullptr (scalar parameter)
    int slot_base[NLS_SP_MAXSLOT];           // first global index of the parameter (vector) slot s refers to
    long long nterms;
    const double *theta;                     // [P] global parameters (device)
    double *tv;                              // [nterms] term values
    double *jv;                              // [k][nterms] partial derivatives
    unsigned long long *nbad;                // [2] += terms whose value / whose derivatives are not finite
};
