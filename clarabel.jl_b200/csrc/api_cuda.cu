// C-ABI implementation of the device-resident KKT solver (include/clarabel_b200.h).
// Host orchestration only: every numeric operation is a kernel from kernels.cuh.  There is no
// CPU fallback: if CUDA is unavailable every entry point returns a negative status.
#include "../../include/clarabel_b200.h"
#include "symbolic.h"
#include "api_common.h"
#include "kernels.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>
#include <string>
#include <vector>

namespace cb200 {

#define CUDA_OK(call)                                                                     \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess) {                                                          \
            set_error(std::string(#call) + ": " + cudaGetErrorString(_e));                \
            return -100 - (int)_e;                                                        \
        }                                                                                 \
    } while (0)

template <class T> struct DevBuf {
    T* p = nullptr; size_t n = 0;
    // New device buffers are zero-filled: cudaMalloc hands back whatever an earlier owner of the pages left
    // there, and a solver created late in a long-lived process must behave like one in a fresh process.
    // CB200_POISON=1 fills with 0xFF instead (NaN doubles, -1 indices: any read-before-write that matters
    // becomes a loud failure; the GPU suite is run once per round that way), CB200_POISON=2 leaves the
    // memory untouched (for compute-sanitizer --tool initcheck).
    static int fill_mode() {
        static const int m = [] { const char* e = getenv("CB200_POISON"); return e ? atoi(e) : 0; }();
        return m;
    }
    cudaError_t alloc(size_t count, bool fill = true) {
        free(); n = count;
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e != cudaSuccess) return e;
        const int m = fill_mode();
        if (m == 2 || !fill) return cudaSuccess;
        e = cudaMemset(p, m == 1 ? 0xFF : 0, count * sizeof(T));
        if (e != cudaSuccess) return e;
        return cudaDeviceSynchronize();      // the fill runs on the NULL stream, the solver's streams do not wait for it
    }
    cudaError_t upload(const std::vector<T>& v, cudaStream_t st = 0) {
        cudaError_t e = alloc(v.size(), /*fill=*/false);
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st);
    }
    void free() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { free(); }
};

struct Batch { int32_t off = 0, cnt = 0, maxnf = 0, maxns = 0, maxnr = 0, maxpanel = 0; bool many_children = false;
               int32_t schur_t128 = 0, schur_t64 = 0;
               int32_t src_per_row = 0;
               bool all_padded = true;
               std::vector<int32_t> ns_desc;  // large batches are sorted by pivot width (descending): the fronts still
                                              // active at pivot column kb are the first active(kb) of the batch
               int active(int kb) const { int c = 0; while (c < (int)ns_desc.size() && ns_desc[c] > kb) ++c; return c; } };     // every front of the batch is on the padded (TMA-addressable) layout     // max over the fronts of (assembly sources / front rows)      // max over the fronts of the number of update-block tiles per side

constexpr int NSMALL = 6;
static const int kSmallNf[NSMALL] = {16, 32, 64, 96, 128, 152};
constexpr int NPANEL = 4;
static const int kPanelDoubles[NPANEL] = {4096, 9000, 16000, 1 << 30};   // smem need classes (doubles)
constexpr int NSOLVE = 6;      // 0: single-column leaves, 1: warp (16 < ns <= 32), 2: CTA per supernode, 3: multi-CTA (big),
                               // 4: 8 lanes (tiny), 5: warp, ns <= 16 (half the registers, twice the warps in flight)

struct LevelPlan {
    Batch small[NSMALL];
    Batch panel[NPANEL];       // 64 < nf <= 152: panel-in-smem kernel, classes by panel size
    Batch large;
    Batch solve[NSOLVE];
    Batch topf, tops;          // multi-GPU: replicated top fronts of this level (factor / solve)
    int64_t wtotal = 0;        // doubles of W workspace needed by the large batch
};

struct Timers {
    enum { CONE = 0, FACTOR = 1, SOLVE = 2, SPMV = 3, SCHUR = 4, PANEL = 5, SMALL = 6, ASM = 7, NCOARSE = 8 };
    // per-kernel-class segments, recorded only in detail mode 2 (cb200_get_fine_timers)
    enum { F_SMALL0 = 0, F_PANEL0 = 6, F_ZERO = 10, F_ASM = 11, F_DIAG = 12, F_ROWS = 13, F_SCHUR = 14, F_FINISH = 15,
           F_PROLOGUE = 16, F_FWD_LEAF = 17, F_FWD_SUB = 18, F_FWD_WARP = 19, F_FWD_CTA = 20, F_FWD_BIG_ASM = 21,
           F_FWD_BIG_TRI = 22, F_FWD_BIG_GEMV = 23, F_BWD_LEAF = 24, F_BWD_SUB = 25, F_BWD_WARP = 26, F_BWD_CTA = 27,
           F_BWD_BIG_GEMV = 28, F_BWD_BIG_TRI = 29, F_PERMUTE = 30, F_COMM = 31, F_PANELUPD = 32, NFINE = 33 };
    static constexpr int NPH = NCOARSE + NFINE;
    double ms[NPH] = {};
    double nfactor = 0, nsolve = 0, nlaunch = 0;
    std::vector<cudaEvent_t> pool;
    struct Seg { int ph; cudaEvent_t a, b; bool closed; };
    std::vector<Seg> open;
    size_t used = 0;
    cudaEvent_t get() {
        if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
        return pool[used++];
    }
    void begin(int ph, cudaStream_t st) { Seg s{ph, get(), get(), false}; cudaEventRecord(s.a, st); open.push_back(s); }
    void end(cudaStream_t st) {        // closes the most recently opened, still open segment (segments nest)
        for (size_t i = open.size(); i-- > 0;) if (!open[i].closed) { cudaEventRecord(open[i].b, st); open[i].closed = true; return; }
    }
    void collect() {      // call after a stream synchronise
        for (auto& s : open) { float t = 0; if (s.closed && cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) ms[s.ph] += t; }
        open.clear(); used = 0;
    }
    ~Timers() { for (auto e : pool) cudaEventDestroy(e); }
};

// NCCL is bound at run time (dlsym): the library loads without NCCL for single-GPU use, and under
// torchrun it resolves to the NCCL that torch already loaded.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    bool load() {
        if (ok) return true;
        void* hd = RTLD_DEFAULT;
        auto sym = [&](const char* n) { return dlsym(hd, n); };
        if (!sym("ncclAllReduce")) {
            const char* names[] = {"libnccl.so.2", "libnccl.so"};
            hd = nullptr;
            for (const char* n : names) { hd = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (hd) break; }
            if (!hd) return false;
        }
        GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
        AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        ok = GetUniqueId && CommInitRank && AllReduce && GroupStart && GroupEnd && CommDestroy;
        return ok;
    }
};
static NcclApi g_nccl;

#define NCCL_OK(call)                                                                      \
    do {                                                                                   \
        ncclResult_t _r = (call);                                                          \
        if (_r != ncclSuccess) {                                                           \
            set_error(std::string(#call) + ": " +                                          \
                      (g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error")); \
            return -200 - (int)_r;                                                         \
        }                                                                                  \
    } while (0)

struct GraphExec {
    cudaGraphExec_t exec = nullptr;
    double nlaunch = 0;
    bool failed = false;
    ~GraphExec() { if (exec) cudaGraphExecDestroy(exec); }
};

}  // namespace cb200

using namespace cb200;

struct cb200_handle {
    cb200_settings st;
    Symbolic S;
    int64_t N = 0, nnzK = 0, base = 0;
    cudaStream_t stream = nullptr;
    // K (original order, upper CSC) + row-wise index for the symmetric product
    DevBuf<int64_t> d_cp; DevBuf<int32_t> d_ri; DevBuf<double> d_nz;
    DevBuf<int64_t> d_tp; DevBuf<int32_t> d_tc; DevBuf<int64_t> d_tpos;
    DevBuf<int64_t> d_amap, d_diagidx;
    DevBuf<int32_t> d_long_rows; int32_t n_long_rows = 0;
    DevBuf<int8_t> d_dsign_perm, d_dsign_orig;
    DevBuf<int32_t> d_perm;
    // symbolic
    DevBuf<int32_t> d_sn_first, d_rows, d_rel, d_child_ptr, d_child_list, d_batches, d_ld;
    DevBuf<SolveDesc> d_sdesc;          // one packed descriptor per batch entry (same indexing as d_batches)
    DevBuf<int64_t> d_rows_ptr, d_panel_off, d_upd_off, d_woff, d_front_ptr, d_asm_base;
    DevBuf<int32_t> d_asm_colptr, d_asm_src, d_asm_child;
    // numeric
    DevBuf<double> d_L, d_U, d_W, d_D, d_Dinv, d_uvec, d_partial;
    DevBuf<double> d_b, d_x, d_e, d_dx, d_y, d_rx, d_rz;
    DevBuf<double> d_eps; DevBuf<unsigned long long> d_scal;   // [0] max|diag|, [1] normb, [2] norme
    DevBuf<unsigned int> d_nreg; DevBuf<int32_t> d_reglog;
    std::vector<LevelPlan> plan;
    std::vector<int32_t> h_batches;
    bool have_diag = false;
    // fused (outer) boundary state
    bool maps_set = false;
    int64_t n = 0, m = 0, p = 0;
    DevBuf<int64_t> d_mapP, d_mapA;
    // diagonal Hs entries
    int64_t ndiag = 0;
    DevBuf<int8_t> d_dg_kind; DevBuf<int32_t> d_dg_midx, d_dg_cone; DevBuf<int64_t> d_dg_map;
    // dense SOC
    int32_t nsocd = 0;
    DevBuf<int32_t> d_sd_moff, d_sd_dim, d_sd_socid; DevBuf<int64_t> d_sd_hoff;
    DevBuf<int64_t> d_mapHs;
    // sparse SOC expansion
    int64_t nexp = 0; int32_t nsocs = 0;
    DevBuf<int32_t> d_ex_src, d_ex_cone, d_exD_cone; DevBuf<int64_t> d_mapu, d_mapv, d_mapD;
    // PSD
    int32_t npsd = 0, psd_maxn = 0;
    DevBuf<int32_t> d_psd_side; DevBuf<int64_t> d_psd_roff, d_psd_hoff;
    DevBuf<double> d_psd_R, d_psd_A;
    int64_t psd_rtotal = 0;
    // cone state
    int64_t nsoc = 0, nsocrows = 0;
    DevBuf<double> d_w, d_eta, d_socd, d_socu, d_socv;
    double last_eps = 0;
    bool resident = false;
    int detail = 0;            // 1: per-phase event timing, 2: also per kernel class (both disable graph replay)
    // size-class boundaries of the factorisation plan (tuning knobs, environment variables
    // CB200_PANEL_MIN_NF / CB200_PANEL_MAX_NF / CB200_SMALL_MAX_NF / CB200_NO_PANEL, read once into the
    // symbolic options because the large path pads its panels): panel kernel for
    // panel_min_nf < nf <= panel_max_nf, shared-memory front kernel up to small_max_nf, pivot-block
    // (large) path above
    SymbolicOptions opt;
    bool to_panel(int nf, int ns) const { return front_to_panel(opt, nf, ns); }
    bool to_large(int nf, int ns) const { return front_is_large(opt, nf, ns); }
    // TMA descriptors of the large-front panels (one CUtensorMap per large front)
    bool use_tma = false; int tma_kmajor = 1;
    int tma_tile = 64;                // CTA tile of the TMA GEMM: 64 (3 CTAs per SM; measured best on every workload:
                                      // C5 Schur 5.4 ms against 8.4 with 128 and 7.5 with the LDG kernel, C4 32 / 42.5 / 42.9),
                                      // 128, or 0 = by front size (CB200_TMA_TILE)
    DevBuf<CUtensorMap> d_tmaps, d_tmaps64; DevBuf<int32_t> d_tmap_of;
    // multi-GPU state
    bool dist = false; int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    std::vector<int32_t> owner; std::vector<int8_t> is_top;
    std::vector<int32_t> h_top_list;
    std::vector<std::vector<int32_t>> h_top_by_level;
    DevBuf<int8_t> d_active, d_keepcol, d_topcolkeep;
    DevBuf<int32_t> d_top_list;
    GraphExec g_factor[2], g_solve;     // CUDA graphs: factor (without/with static reg), solve sweeps
    // Fork/join streams: the kernel classes of one tree level are independent of each other (small
    // fronts vs the pivot-block chain of the large fronts; leaf / warp / CTA / big solve classes),
    // most of them are latency-bound and leave SMs idle, so they run as parallel branches (captured
    // as such into the CUDA graphs).  CB200_MULTISTREAM=0 serialises them again.
    cudaStream_t side[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
    bool multi_stream = true;
    bool sort_batches = true;          // CB200_SORT_BATCHES=0: keep the size order of the level lists
    bool merged_rows = true;           // 16-column warp solve kernels fetch the triangle and the first L21 chunk in ONE batch
                                       // of loads (one dependent round trip less; 80 registers, 3 CTAs/SM).  Measured on
                                       // C5 / C3: fwd -8 % / -5 %, bwd -10 % / -7 %.  CB200_MERGED_ROWS=0: two-phase loads.
    // bottom subtrees of narrow supernodes solved by one CTA each (k_fwd_subtree / k_bwd_subtree)
    bool use_subtree = false;          // CB200_SUBTREE=1 turns them on.  Measured (C5): 5.58 ms per solve against 3.19 ms
                                       // level-scheduled - with only L streamed from HBM but metadata still fetched per
                                       // supernode, 16 warps per SM are too few to hide the dependent round trips.  Kept as
                                       // the base of a version that bulk-prefetches the (contiguous) subtree panels.
    int32_t nsub = 0, sub_nlev = 0, sub_maxcols = 0, sub_maxrows = 0;
    DevBuf<int32_t> d_sub_sn0, d_sub_sn1, d_sub_order, d_sub_lvl;
    DevBuf<int64_t> d_sub_optr;
    bool ms_on() const { return multi_stream && detail == 0; }
    void fork(int n) {            // side[0..n) start after everything issued so far on `stream`
        cudaEventRecord(ev_fork, stream);
        for (int i = 0; i < n; ++i) cudaStreamWaitEvent(side[i], ev_fork, 0);
    }
    void join(int n) {            // `stream` continues after side[0..n)
        for (int i = 0; i < n; ++i) { cudaEventRecord(ev_join[i], side[i]); cudaStreamWaitEvent(stream, ev_join[i], 0); }
    }
    // grow-only device staging for the inner-boundary uploads (update_values!/scale_values!/update_P/A):
    // no cudaMalloc / cudaFree per call and no stream sync for pageable caller buffers
    DevBuf<int64_t> d_stage_idx; DevBuf<double> d_stage_val;
    size_t stage_idx_cap = 0, stage_val_cap = 0;
    Timers tm;
};

namespace {

DevSym devsym(cb200_handle* h) {
    DevSym d;
    d.sn_first = h->d_sn_first.p; d.rows_ptr = h->d_rows_ptr.p; d.rows = h->d_rows.p; d.rel = h->d_rel.p;
    d.child_ptr = h->d_child_ptr.p; d.child_list = h->d_child_list.p;
    d.panel_off = h->d_panel_off.p; d.upd_off = h->d_upd_off.p; d.dsign = h->d_dsign_perm.p; d.ld = h->d_ld.p;
    d.front_ptr = h->d_front_ptr.p; d.asm_base = h->d_asm_base.p; d.asm_colptr = h->d_asm_colptr.p;
    d.asm_src = h->d_asm_src.p; d.asm_child = h->d_asm_child.p;
    d.active = h->dist ? h->d_active.p : nullptr;
    return d;
}

inline int nblk(int64_t n, int t) { return (int)((n + t - 1) / t); }

#define LAUNCH(h) ((h)->tm.nlaunch += 1)

// times the launches issued while it is alive under one kernel class (detail mode 2 only)
struct FineScope {
    cb200_handle* h; bool on;
    FineScope(cb200_handle* h_, int cls) : h(h_), on(h_->detail >= 2) { if (on) h->tm.begin(Timers::NCOARSE + cls, h->stream); }
    ~FineScope() { if (on) h->tm.end(h->stream); }
};

template <int T>
int launch_small(cb200_handle* h, const Batch& b, int cls, RegParams rp, cudaStream_t st) {
    if (b.cnt == 0) return 0;
    FineScope fs(h, Timers::F_SMALL0 + cls);
    const int sbuf = std::min(SB, b.maxns);
    size_t sm = ((size_t)b.maxnf * b.maxnf + (size_t)sbuf * sbuf) * sizeof(double);
    k_factor_small<T><<<b.cnt, T, sm, st>>>(devsym(h), h->d_batches.p + b.off, h->d_L.p,
                                                   h->d_U.p, h->d_D.p, h->d_Dinv.p, rp, h->d_nreg.p);
    LAUNCH(h);
    return 0;
}

void launch_fwd_level(cb200_handle* h, const LevelPlan& P, int lv) {
    DevSym ds = devsym(h);
    // classes of a level are independent: leaf + tiny on side[0], warp on side[1], CTA on side[2], big on the
    // main stream (fork/join only when a second class has work)
    const bool big_work = P.solve[3].cnt || P.tops.cnt;
    const int nclass = (P.solve[0].cnt || P.solve[4].cnt) + (P.solve[1].cnt != 0 || P.solve[5].cnt != 0) + (P.solve[2].cnt != 0) + big_work;
    const bool ms = h->ms_on() && nclass > 1;
    cudaStream_t s_leaf = ms ? h->side[0] : h->stream, s_warp = ms ? h->side[1] : h->stream,
                 s_cta = ms ? h->side[2] : h->stream;
    if (ms) h->fork(3);
    const Batch& b0 = P.solve[0];
    if (b0.cnt) {
        { FineScope fs(h, Timers::F_FWD_LEAF);
        k_fwd_leaf<<<nblk(b0.cnt, 128), 128, 0, s_leaf>>>(ds, h->d_batches.p + b0.off, b0.cnt, h->d_L.p,
                                                             h->d_y.p, h->d_uvec.p);
        LAUNCH(h);
        }
    }
    const Batch& b1 = P.solve[1];
    if (b1.cnt) {
        FineScope fs(h, Timers::F_FWD_WARP);
        k_fwd_warp<32, false><<<nblk(b1.cnt, WPB), WPB * 32, (size_t)WPB * b1.maxnf * sizeof(double), s_warp>>>(
            ds, h->d_sdesc.p + b1.off, b1.cnt, b1.maxnf, h->d_L.p, h->d_y.p, h->d_uvec.p);
        LAUNCH(h);
    }
    const Batch& b5 = P.solve[5];
    if (b5.cnt) {
        FineScope fs(h, Timers::F_FWD_WARP);
        if (h->merged_rows) k_fwd_warp<16, true><<<nblk(b5.cnt, WPB), WPB * 32, (size_t)WPB * b5.maxnf * sizeof(double), s_warp>>>(
            ds, h->d_sdesc.p + b5.off, b5.cnt, b5.maxnf, h->d_L.p, h->d_y.p, h->d_uvec.p);
        else k_fwd_warp<16, false><<<nblk(b5.cnt, WPB), WPB * 32, (size_t)WPB * b5.maxnf * sizeof(double), s_warp>>>(
            ds, h->d_sdesc.p + b5.off, b5.cnt, b5.maxnf, h->d_L.p, h->d_y.p, h->d_uvec.p);
        LAUNCH(h);
    }
    const Batch& b2 = P.solve[2];
    if (b2.cnt) {
        { FineScope fs(h, Timers::F_FWD_CTA);
        k_fwd_cta<<<b2.cnt, 256, (size_t)b2.maxnf * sizeof(double), s_cta>>>(
            ds, h->d_sdesc.p + b2.off, h->d_L.p, h->d_y.p, h->d_uvec.p);
        LAUNCH(h);
        }
    }
    for (int pass = 0; pass < 2; ++pass) {
        const Batch& b3 = pass == 0 ? P.solve[3] : P.tops;
        if (!b3.cnt) continue;
        const int32_t* bl = h->d_batches.p + b3.off;
        { FineScope fs(h, Timers::F_FWD_BIG_ASM);
        if (b3.src_per_row >= 512)
            k_big_asm_fwd<256><<<dim3(b3.maxnf, b3.cnt), 256, 0, h->stream>>>(ds, bl, h->d_y.p, h->d_uvec.p);
        else if (b3.many_children)
            k_big_asm_fwd<32><<<dim3(nblk((int64_t)b3.maxnf * 32, 256), b3.cnt), 256, 0, h->stream>>>(ds, bl, h->d_y.p, h->d_uvec.p);
        else
            k_big_asm_fwd<1><<<dim3(nblk(b3.maxnf, 256), b3.cnt), 256, 0, h->stream>>>(ds, bl, h->d_y.p, h->d_uvec.p);
        LAUNCH(h);
        }
        if (pass == 1 && h->dist) {
            // sum the per-rank partial right-hand sides of the replicated top fronts
            FineScope fs(h, Timers::F_COMM);
            const Symbolic& S = h->S;
            g_nccl.GroupStart();
            for (int32_t sn : h->h_top_by_level[lv]) {
                double* yt = h->d_y.p + S.sn_first[sn];
                g_nccl.AllReduce(yt, yt, (size_t)S.ns(sn), ncclDouble, ncclSum, h->comm, h->stream);
                if (S.nr(sn) > 0) {
                    double* ut = h->d_uvec.p + S.rows_ptr[sn];
                    g_nccl.AllReduce(ut, ut, (size_t)S.nr(sn), ncclDouble, ncclSum, h->comm, h->stream);
                }
            }
            g_nccl.GroupEnd();
        }
        const int npanel = nblk(b3.maxns, WP);
        for (int pk = 0; pk < npanel; ++pk) {
            { FineScope fs(h, Timers::F_FWD_BIG_TRI);
              k_big_tri_fwd<<<b3.cnt, 256, 0, h->stream>>>(ds, bl, pk, h->d_L.p, h->d_y.p);
              LAUNCH(h); }
            const int rows_below = b3.maxnf - pk * WP;     // upper bound
            if (rows_below > 0) {
                FineScope fs(h, Timers::F_FWD_BIG_GEMV);
                k_big_gemv_fwd<<<dim3(nblk(rows_below, BRT), b3.cnt), 256, 0, h->stream>>>(
                    ds, bl, pk, h->d_L.p, h->d_y.p, h->d_uvec.p);
                LAUNCH(h);
            }
        }
    }
    if (ms) h->join(3);
}
void launch_bwd_level(cb200_handle* h, const LevelPlan& P) {
    DevSym ds = devsym(h);
    const bool big_work = P.solve[3].cnt || P.tops.cnt;
    const int nclass = (P.solve[0].cnt || P.solve[4].cnt) + (P.solve[1].cnt != 0 || P.solve[5].cnt != 0) + (P.solve[2].cnt != 0) + big_work;
    const bool ms = h->ms_on() && nclass > 1;
    cudaStream_t s_leaf = ms ? h->side[0] : h->stream, s_warp = ms ? h->side[1] : h->stream,
                 s_cta = ms ? h->side[2] : h->stream;
    if (ms) h->fork(3);
    for (int pass = 0; pass < 2; ++pass) {
        const Batch& b3 = pass == 0 ? P.tops : P.solve[3];
        if (!b3.cnt) continue;
        const int32_t* bl = h->d_batches.p + b3.off;
        const int npanel = nblk(b3.maxns, WP);
        const int maxtiles = std::max(1, nblk(b3.maxnf, BRT));
        for (int pk = npanel - 1; pk >= 0; --pk) {
            const int rows_below = b3.maxnf - pk * WP;
            if (rows_below > 0) {
                FineScope fs(h, Timers::F_BWD_BIG_GEMV);
                k_big_gemvT_bwd<<<dim3(nblk(rows_below, BRT), b3.cnt), 256, 0, h->stream>>>(
                    ds, bl, pk, maxtiles, h->d_L.p, h->d_y.p, h->d_partial.p);
                LAUNCH(h);
            }
            FineScope fs(h, Timers::F_BWD_BIG_TRI);
            k_big_tri_bwd<<<b3.cnt, 256, 0, h->stream>>>(ds, bl, pk, maxtiles, h->d_L.p, h->d_Dinv.p,
                                                         h->d_partial.p, h->d_y.p);
            LAUNCH(h);
        }
    }
    const Batch& b2 = P.solve[2];
    if (b2.cnt) {
        { FineScope fs(h, Timers::F_BWD_CTA);
        k_bwd_cta<<<b2.cnt, 256, (size_t)b2.maxnf * sizeof(double), s_cta>>>(
            ds, h->d_batches.p + b2.off, h->d_L.p, h->d_Dinv.p, h->d_y.p);
        LAUNCH(h);
        }
    }
    const Batch& b1 = P.solve[1];
    if (b1.cnt) {
        FineScope fs(h, Timers::F_BWD_WARP);
        k_bwd_warp<32, false><<<nblk(b1.cnt, WPB), WPB * 32, (size_t)WPB * (b1.maxnf + 32 + 32 * BT_LD) * sizeof(double), s_warp>>>(
            ds, h->d_sdesc.p + b1.off, b1.cnt, b1.maxnf, h->d_L.p, h->d_Dinv.p, h->d_y.p);
        LAUNCH(h);
    }
    const Batch& b5 = P.solve[5];
    if (b5.cnt) {
        FineScope fs(h, Timers::F_BWD_WARP);
        if (h->merged_rows) k_bwd_warp<16, true><<<nblk(b5.cnt, WPB), WPB * 32, (size_t)WPB * (b5.maxnf + 32 + 32 * 17) * sizeof(double), s_warp>>>(
            ds, h->d_sdesc.p + b5.off, b5.cnt, b5.maxnf, h->d_L.p, h->d_Dinv.p, h->d_y.p);
        else k_bwd_warp<16, false><<<nblk(b5.cnt, WPB), WPB * 32, (size_t)WPB * (b5.maxnf + 32 + 32 * 17) * sizeof(double), s_warp>>>(
            ds, h->d_sdesc.p + b5.off, b5.cnt, b5.maxnf, h->d_L.p, h->d_Dinv.p, h->d_y.p);
        LAUNCH(h);
    }
    const Batch& b0 = P.solve[0];
    if (b0.cnt) {
        { FineScope fs(h, Timers::F_BWD_LEAF);
        k_bwd_leaf<<<nblk(b0.cnt, 128), 128, 0, s_leaf>>>(ds, h->d_batches.p + b0.off, b0.cnt, h->d_L.p,
                                                             h->d_Dinv.p, h->d_y.p);
        LAUNCH(h);
        }
    }
    if (ms) h->join(3);
}


// One CUtensorMap per large front: the nf x ns panel (leading dimension ld, a multiple of 8) viewed
// as a 3-D tensor so that one TMA box is a 128-row x 16-k operand tile in a DMMA-friendly layout
// (see k_ldl_update_tma).  The encoder lives in libcuda; it is fetched through the runtime
// (cudaGetDriverEntryPoint) so that the library has no link-time dependency on the driver.
// CB200_NO_TMA=1 selects the LDG fallback kernel.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int build_tensor_maps(cb200_handle* h) {
    const Symbolic& S = h->S;
    h->use_tma = false;
    if (const char* e = getenv("CB200_NO_TMA")) if (e[0] == '1') return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !fn) { cudaGetLastError(); return 0; }
    EncodeTiledFn encode = (EncodeTiledFn)fn;
    if (const char* e = getenv("CB200_TMA_TILE")) h->tma_tile = atoi(e);
    std::vector<int32_t> map_of(S.nsuper, -1);
    std::vector<CUtensorMap> maps, maps64;
    int kmajor = 1;
    if (const char* e = getenv("CB200_TMA_NATURAL")) if (e[0] == '1') kmajor = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        maps.clear(); maps64.clear(); std::fill(map_of.begin(), map_of.end(), -1);
        bool ok = true;
        for (int32_t sn = 0; sn < S.nsuper && ok; ++sn) {
            const int ns = S.ns(sn), nf = ns + S.nr(sn);
            if (!h->to_large(nf, ns)) continue;
            const cuuint64_t ld = (cuuint64_t)S.panel_ld[sn];
            CUtensorMap m;
            cuuint64_t dims[3], strides[2]; cuuint32_t box[3], estr[3] = {1, 1, 1};
            if (kmajor) { dims[0] = 8; dims[1] = (cuuint64_t)ns; dims[2] = ld / 8; strides[0] = ld * 8; strides[1] = 64;
                          box[0] = 8; box[1] = TK; box[2] = TB / 8; }
            else        { dims[0] = 8; dims[1] = ld / 8; dims[2] = (cuuint64_t)ns; strides[0] = 64; strides[1] = ld * 8;
                          box[0] = 8; box[1] = TB / 8; box[2] = TK; }
            CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)(h->d_L.p + S.panel_off[sn]), dims, strides,
                                box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { ok = false; break; }
            CUtensorMap m64;                              // same view, box of 64 rows for the small tile
            if (kmajor) box[2] = 64 / 8; else box[1] = 64 / 8;
            r = encode(&m64, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)(h->d_L.p + S.panel_off[sn]), dims, strides,
                       box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { ok = false; break; }
            map_of[sn] = (int32_t)maps.size(); maps.push_back(m); maps64.push_back(m64);
        }
        if (ok) { h->use_tma = true; h->tma_kmajor = kmajor; break; }
        if (kmajor == 0) break;
        kmajor = 0;                              // retry with the natural dimension order
    }
    if (!h->use_tma) return 0;
    if (maps.empty()) { maps.resize(1); maps64.resize(1); }
    CUDA_OK(h->d_tmaps.alloc(maps.size())); CUDA_OK(h->d_tmaps64.alloc(maps64.size()));
    CUDA_OK(cudaMemcpyAsync(h->d_tmaps.p, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaMemcpyAsync(h->d_tmaps64.p, maps64.data(), maps64.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(h->d_tmap_of.upload(map_of, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}

// Builds the per-level launch plans.  In multi-GPU mode a rank only schedules the supernodes it
// owns plus the replicated top fronts (which always take the large-front / big-solve paths so
// that assembly, all-reduce and factorisation are separate steps).
int build_plans(cb200_handle* h) {
    const Symbolic& S = h->S;
    cudaStream_t s = h->stream;
    h->plan.assign(S.nlevels, LevelPlan());
    std::vector<int32_t> batches; std::vector<int64_t> woff;
    int64_t wmax = 0;
    auto add_batch = [&](Batch& b, std::vector<int32_t> v, bool large, LevelPlan& P) {
        b = Batch();
        if (large) {
            std::stable_sort(v.begin(), v.end(), [&](int32_t x, int32_t y) { return S.ns(x) > S.ns(y); });
            for (int32_t sn : v) b.ns_desc.push_back(S.ns(sn));
        } else if (h->sort_batches) {
            // memory order: panels, update blocks and solution slices of consecutive supernodes are
            // adjacent in HBM (postorder layout), so neighbouring warps / CTAs touch neighbouring DRAM
            // pages instead of random ones (the level lists come sorted by front size)
            std::sort(v.begin(), v.end());
        }
        b.off = (int32_t)batches.size(); b.cnt = (int32_t)v.size();
        int64_t w = 0;
        for (int32_t sn : v) {
            int nf = S.ns(sn) + S.nr(sn);
            b.maxnf = std::max(b.maxnf, nf); b.maxns = std::max(b.maxns, S.ns(sn));
            b.maxnr = std::max(b.maxnr, S.nr(sn));
            b.maxpanel = std::max(b.maxpanel, nf * S.ns(sn) + 25 * nf);      // smem need of k_factor_panel
            if (S.child_ptr[sn + 1] - S.child_ptr[sn] > 64) b.many_children = true;
            b.src_per_row = std::max<int32_t>(b.src_per_row, (int32_t)((S.asm_base[sn + 1] - S.asm_base[sn]) / std::max(1, nf)));
            if (large && !h->to_large(nf, S.ns(sn))) b.all_padded = false;   // multi-GPU: small top fronts on the large path
            if (large && S.nr(sn) > 0) {
                b.schur_t128 = std::max(b.schur_t128, (nf - 1) / TB - S.ns(sn) / TB + 1);
                b.schur_t64 = std::max(b.schur_t64, (nf - 1) / GBM - S.ns(sn) / GBM + 1);
            }
            batches.push_back(sn);
            woff.push_back(large ? w : 0);
            if (large) w += (int64_t)nblk(S.ns(sn), PB) * PB * PB;   // parked diagonal blocks
        }
        if (large) { P.wtotal += w; wmax = std::max(wmax, w); }
    };
    // ---- bottom subtrees for the solves: complete subtrees whose supernodes are all narrow (ns <= 32,
    // nf <= SUB_MAXNF) and whose columns / contribution rows fit a shared-memory window
    std::vector<int8_t> in_sub(S.nsuper, 0);
    h->nsub = 0;
    if (h->use_subtree && S.nsuper > 0) {
        constexpr int64_t MAXCOLS = 2048, MAXROWS = 6144, MAXPANEL = 131072;
        constexpr int32_t MINSN = 16;
        const int32_t n = S.nsuper;
        std::vector<int8_t> ok(n, 1);
        std::vector<int64_t> wc(n), wr(n), wp(n); std::vector<int32_t> cnt(n, 1);
        for (int32_t sn = 0; sn < n; ++sn) {            // children precede parents
            const int64_t ns = S.ns(sn), nr = S.nr(sn);
            wc[sn] += ns; wr[sn] += nr; wp[sn] += (ns + nr) * ns;
            bool self = ns <= 32 && ns + nr <= SUB_MAXNF;
            if (h->dist) self = self && !h->is_top[sn] && h->owner[sn] == h->rank;
            if (!self || wc[sn] > MAXCOLS || wr[sn] > MAXROWS || wp[sn] > MAXPANEL) ok[sn] = 0;
            const int32_t p = S.sn_parent[sn];
            if (p >= 0) { wc[p] += wc[sn]; wr[p] += wr[sn]; wp[p] += wp[sn]; cnt[p] += cnt[sn]; if (!ok[sn]) ok[p] = 0; }
        }
        std::vector<int32_t> sn0, sn1, order, lvl; std::vector<int64_t> optr(1, 0);
        int32_t maxl = 0;
        std::vector<int32_t> roots;
        for (int32_t sn = 0; sn < n; ++sn) {
            const int32_t p = S.sn_parent[sn];
            if (ok[sn] && (p < 0 || !ok[p]) && cnt[sn] >= MINSN) { roots.push_back(sn); maxl = std::max(maxl, S.sn_level[sn] + 1); }
        }
        h->sub_nlev = maxl; h->sub_maxcols = 1; h->sub_maxrows = 1;
        for (int32_t r : roots) {
            const int32_t a = r - cnt[r] + 1;
            sn0.push_back(a); sn1.push_back(r);
            h->sub_maxcols = std::max<int32_t>(h->sub_maxcols, (int32_t)wc[r]);
            h->sub_maxrows = std::max<int32_t>(h->sub_maxrows, (int32_t)wr[r]);
            std::vector<int32_t> off(maxl + 1, 0);
            for (int32_t q = a; q <= r; ++q) { in_sub[q] = 1; off[S.sn_level[q] + 1]++; }
            for (int32_t l = 0; l < maxl; ++l) off[l + 1] += off[l];
            const size_t base = order.size();
            order.resize(base + (size_t)cnt[r]);
            { std::vector<int32_t> pos(off.begin(), off.end() - 1);
              for (int32_t q = a; q <= r; ++q) order[base + pos[S.sn_level[q]]++] = q; }
            lvl.insert(lvl.end(), off.begin(), off.end());
            optr.push_back((int64_t)order.size());
        }
        h->nsub = (int32_t)roots.size();
        if (h->nsub) {
            CUDA_OK(h->d_sub_sn0.upload(sn0, s)); CUDA_OK(h->d_sub_sn1.upload(sn1, s));
            CUDA_OK(h->d_sub_order.upload(order, s)); CUDA_OK(h->d_sub_lvl.upload(lvl, s));
            CUDA_OK(h->d_sub_optr.upload(optr, s));
        }
    }
    for (int lv = 0; lv < S.nlevels; ++lv) {
        std::vector<int32_t> cls[NSMALL + 1], scl[NSOLVE], pcl[NPANEL], top;
        for (int32_t q = S.level_ptr[lv]; q < S.level_ptr[lv + 1]; ++q) {
            int32_t sn = S.level_list[q];
            if (h->dist) {
                if (h->is_top[sn]) { top.push_back(sn); continue; }
                if (h->owner[sn] != h->rank) continue;
            }
            int nf = S.ns(sn) + S.nr(sn);
            int c = 0; while (c < NSMALL && nf > kSmallNf[c]) ++c;
            if (h->to_large(nf, S.ns(sn)) || nf > kSmallNf[NSMALL - 1]) c = NSMALL;
            if (h->to_panel(nf, S.ns(sn))) {
                int pc = 0; while (nf * S.ns(sn) + 25 * nf > kPanelDoubles[pc]) ++pc;
                pcl[pc].push_back(sn);
            } else cls[c].push_back(sn);
            const bool leaf = (S.ns(sn) == 1 && S.child_ptr[sn + 1] == S.child_ptr[sn]);
            const bool big = (int64_t)nf * S.ns(sn) >= 65536 && S.ns(sn) > 32;
            // (the 8-lane kernels for tiny supernodes are gone: the 16-column warp kernel takes them)
            static const bool warp16 = !(getenv("CB200_WARP16") && getenv("CB200_WARP16")[0] == '0');    // debugging aid
            const int d = leaf ? 0 : ((S.ns(sn) <= 32 && nf <= 192) ? (S.ns(sn) <= 16 && warp16 ? 5 : 1) : (big ? 3 : 2));
            if (!in_sub[sn]) scl[d].push_back(sn);
        }
        LevelPlan& P = h->plan[lv];
        P.wtotal = 0;
        for (int c = 0; c < NSMALL; ++c) add_batch(P.small[c], cls[c], false, P);
        for (int c = 0; c < NPANEL; ++c) add_batch(P.panel[c], pcl[c], false, P);
        add_batch(P.large, cls[NSMALL], true, P);
        for (int d = 0; d < NSOLVE; ++d) add_batch(P.solve[d], scl[d], false, P);
        add_batch(P.topf, top, true, P);
        add_batch(P.tops, top, false, P);
    }
    h->h_batches = batches;
    CUDA_OK(h->d_batches.upload(batches, s)); CUDA_OK(h->d_woff.upload(woff, s));
    {
        std::vector<SolveDesc> sd(batches.size());
        for (size_t i = 0; i < batches.size(); ++i) {
            const int32_t sn = batches[i];
            SolveDesc& d = sd[i];
            d.f = S.sn_first[sn]; d.ns = S.ns(sn); d.nr = S.nr(sn); d.ld = S.panel_ld[sn];
            d.nchild = S.child_ptr[sn + 1] - S.child_ptr[sn]; d.pad = 0; d.pad2 = 0;
            d.panel_off = S.panel_off[sn]; d.rows_ptr = S.rows_ptr[sn]; d.front_ptr = S.front_ptr[sn]; d.asm_base = S.asm_base[sn];
        }
        CUDA_OK(h->d_sdesc.alloc(std::max<size_t>(1, sd.size())));
        if (!sd.empty()) CUDA_OK(cudaMemcpyAsync(h->d_sdesc.p, sd.data(), sd.size() * sizeof(SolveDesc), cudaMemcpyHostToDevice, s));
        CUDA_OK(cudaStreamSynchronize(s));
    }
    CUDA_OK(h->d_W.alloc((size_t)std::max<int64_t>(1, wmax)));
    size_t pmax = 1;
    for (const LevelPlan& P : h->plan)
        for (const Batch* b3 : {&P.solve[3], &P.tops})
            if (b3->cnt) pmax = std::max(pmax, (size_t)b3->cnt * std::max(1, nblk(b3->maxnf, BRT)) * WP);
    CUDA_OK(h->d_partial.alloc(pmax));
    return 0;
}

// all-reduce (sum) of the panel and update block of every front of a batch, one NCCL group
int allreduce_fronts(cb200_handle* h, const Batch& B, const std::vector<int32_t>& list) {
    const Symbolic& S = h->S;
    NCCL_OK(g_nccl.GroupStart());
    for (int32_t k = 0; k < B.cnt; ++k) {
        const int32_t sn = list[k];
        const int64_t ns = S.ns(sn), nr = S.nr(sn), ld = S.panel_ld[sn];
        double* Lp = h->d_L.p + S.panel_off[sn];
        NCCL_OK(g_nccl.AllReduce(Lp, Lp, (size_t)(ld * ns), ncclDouble, ncclSum, h->comm, h->stream));
        if (nr > 0) {
            double* Us = h->d_U.p + S.upd_off[sn];
            NCCL_OK(g_nccl.AllReduce(Us, Us, (size_t)(nr * nr), ncclDouble, ncclSum, h->comm, h->stream));
        }
    }
    NCCL_OK(g_nccl.GroupEnd());
    return 0;
}

// zero the update blocks of the large fronts of one level
__global__ void k_zero_upd(DevSym S, const int32_t* batch, double* Ust) {
    const int s = batch[blockIdx.y];
    const int64_t nr = S.rows_ptr[s + 1] - S.rows_ptr[s];
    double* U = Ust + S.upd_off[s];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nr * nr;
         i += (int64_t)gridDim.x * blockDim.x) U[i] = 0.0;
}

// numeric factorisation of whatever is in d_nz (+ optional on-device static regularisation)
// Replays `body` (a fixed sequence of stream operations on h->stream) through a CUDA graph: the
// launch sequence of a factorisation / solve sweep only depends on the symbolic structure, so it
// is captured once and relaunched with one call per IP iteration.
template <class F> int run_captured(cb200_handle* h, GraphExec& g, F body, int bit = 1) {
    // use_cuda_graph is a bit mask: 1 = solve sweeps, 2 = factorisation
    if (!(h->st.use_cuda_graph & bit) || g.failed || h->detail) return body();
    if (!g.exec) {
        const double l0 = h->tm.nlaunch;
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError(); g.failed = true; return body();
        }
        int rc = body();
        cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
        g.nlaunch = h->tm.nlaunch - l0; h->tm.nlaunch = l0;
        if (rc != 0 || e != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError(); g.failed = true;
            return rc != 0 ? rc : body();
        }
        e = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { cudaGetLastError(); g.exec = nullptr; g.failed = true; return body(); }
    }
    CUDA_OK(cudaGraphLaunch(g.exec, h->stream));
    h->tm.nlaunch += g.nlaunch;
    return 0;
}

int factor_body(cb200_handle* h, bool static_reg) {
    cudaStream_t st = h->stream;
    const auto& S = h->S;
    if (h->detail >= 2) h->tm.begin(Timers::NCOARSE + Timers::F_PROLOGUE, st);
    CUDA_OK(cudaMemsetAsync(h->d_L.p, 0, h->d_L.n * sizeof(double), st));
    CUDA_OK(cudaMemsetAsync(h->d_nreg.p, 0, sizeof(unsigned int), st));
    if (h->nnzK) { k_scatter<<<nblk(h->nnzK, 256), 256, 0, st>>>(h->d_nz.p, h->d_amap.p, h->nnzK, h->d_L.p); LAUNCH(h); }
    if (static_reg && h->st.static_regularization_enable && h->N) {
        CUDA_OK(cudaMemsetAsync(h->d_scal.p, 0, sizeof(unsigned long long), st));
        k_diag_absmax<<<std::min(nblk(h->N, 256), 1184), 256, 0, st>>>(h->d_nz.p, h->d_diagidx.p, h->N, h->d_scal.p);
        k_compute_eps<<<1, 1, 0, st>>>(h->d_scal.p, h->st.static_regularization_constant,
                                       h->st.static_regularization_proportional, h->d_eps.p);
        k_shift_diag<<<nblk(h->N, 256), 256, 0, st>>>(h->d_nz.p, h->d_diagidx.p, h->d_amap.p,
                                                      h->d_dsign_orig.p, h->d_eps.p, h->N, h->d_L.p);
        h->tm.nlaunch += 3;
    }
    RegParams rp{h->st.dynamic_regularization_eps, h->st.dynamic_regularization_delta,
                 h->st.dynamic_regularization_enable, h->d_reglog.p};
    DevSym ds = devsym(h);
    if (h->dist && h->rank != 0 && !h->h_top_list.empty()) {
        // the original entries of the replicated top fronts are contributed by rank 0 only
        k_zero_panels<<<dim3(32, (unsigned)h->h_top_list.size()), 256, 0, st>>>(ds, h->d_top_list.p, h->d_L.p);
        LAUNCH(h);
    }
    if (h->detail >= 2) h->tm.end(st);
    for (int lv = 0; lv < S.nlevels; ++lv) {
        const LevelPlan& P = h->plan[lv];
        bool small_work = false;
        for (int c = 0; c < NSMALL; ++c) small_work |= P.small[c].cnt != 0;
        for (int c = 0; c < NPANEL; ++c) small_work |= P.panel[c].cnt != 0;
        // the small-front kernels of a level run beside the pivot-block chain of its large fronts
        const bool ms = h->ms_on() && small_work && (P.large.cnt || P.topf.cnt);
        cudaStream_t ss = ms ? h->side[0] : st;
        if (ms) h->fork(1);
        if (h->detail) h->tm.begin(Timers::SMALL, st);
        launch_small<32>(h, P.small[0], 0, rp, ss);        // size classes kSmallNf[0..5]
        launch_small<64>(h, P.small[1], 1, rp, ss);
        launch_small<128>(h, P.small[2], 2, rp, ss);
        launch_small<256>(h, P.small[3], 3, rp, ss);
        launch_small<256>(h, P.small[4], 4, rp, ss);
        launch_small<256>(h, P.small[5], 5, rp, ss);
        for (int c = 0; c < NPANEL; ++c) {
            const Batch& b = P.panel[c];
            if (!b.cnt) continue;
            const size_t sm = (size_t)b.maxpanel * sizeof(double);
            FineScope fs(h, Timers::F_PANEL0 + c);
            k_factor_panel<<<b.cnt, 256, sm, ss>>>(ds, h->d_batches.p + b.off, b.maxpanel, b.maxnf, h->d_L.p, h->d_U.p,
                                                   h->d_D.p, h->d_Dinv.p, rp, h->d_nreg.p);
            LAUNCH(h);
        }
        if (h->detail) h->tm.end(st);
        for (int pass = 0; pass < 2; ++pass) {
            const Batch& B = pass == 0 ? P.large : P.topf;
            if (!B.cnt) continue;
            const bool is_top = pass == 1;
            const int32_t* bl = h->d_batches.p + B.off;
            const int64_t* wo = h->d_woff.p + B.off;
            if (h->detail) h->tm.begin(Timers::ASM, st);
            { FineScope fs(h, Timers::F_ZERO);
              k_zero_upd<<<dim3(std::max(1, std::min(64, nblk((int64_t)B.maxnr * B.maxnr, 1024))), B.cnt), 256, 0, st>>>(ds, bl, h->d_U.p);
              LAUNCH(h); }
            { FineScope fs(h, Timers::F_ASM);
              k_assemble_large<<<dim3(nblk(B.maxnf, ASM_CW), B.cnt), 256, 0, st>>>(ds, bl, h->d_L.p, h->d_U.p);
              LAUNCH(h);
              for (int32_t k = 0; k < B.cnt; ++k) {
                  const int32_t sn = h->h_batches[B.off + k];
                  const int32_t nch = S.child_ptr[sn + 1] - S.child_ptr[sn];
                  if (nch > MANY_CHILDREN) { k_assemble_atomic<<<nch, 64, 0, st>>>(ds, sn, h->d_L.p, h->d_U.p); LAUNCH(h); }
              } }
            if (is_top && h->dist) {
                // root-front assembly across GPUs: every rank holds the contributions of its own
                // subtrees (rank 0 also the original entries); sum them over NVLink.
                FineScope fs(h, Timers::F_COMM);
                std::vector<int32_t> list(h->h_top_by_level[lv]);
                int rc = allreduce_fronts(h, B, list);
                if (rc) return rc;
            }
            if (h->detail) { h->tm.end(st); h->tm.begin(Timers::PANEL, st); }
            const size_t smrows = (size_t)(2 * PB * (PB + 1)) * sizeof(double);
            // one update launch (either version): mode 0 = panel step at pivot block J0, mode 1 = Schur
            auto launch_update = [&](int mode, int J0) {
                const bool tma = h->use_tma && B.all_padded;
                const int T = !tma ? GBM : (h->tma_tile == 64 || h->tma_tile == 128 ? h->tma_tile : (B.maxnf >= 1024 ? 128 : 64));
                int gx, na = 1;
                if (mode == 0) {
                    const int tj0 = (J0 + PB) / T;
                    const int nbt = (B.maxns - 1) / T - tj0 + 1;          // panel column tiles still to update
                    na = (B.maxnf - 1) / T - tj0 + 1;
                    if (nbt <= 0 || na <= 0) return;
                    gx = nbt * na;
                } else {
                    const int t = T == 128 ? B.schur_t128 : B.schur_t64;
                    if (t <= 0) return;
                    gx = t * (t + 1) / 2;
                }
                const int ny = mode == 0 ? B.active(J0 + PB) : B.cnt;     // fronts that still have panel columns to update
                if (ny <= 0) return;
                static const int dbg_fence = getenv("CB200_TMA_FENCE") ? 2 : 0;                       // debugging aids
                static const size_t dbg_pad = getenv("CB200_TMA_PAD_KB") ? (size_t)atoi(getenv("CB200_TMA_PAD_KB")) * 1024 : 0;
                if (tma && T == 128)
                    k_ldl_update_tma<128><<<dim3(gx, ny), tma_threads(128), tma_gemm_smem(128), st>>>(
                        ds, bl, h->d_tmaps.p, h->d_tmap_of.p, mode, J0, na, h->tma_kmajor | dbg_fence, h->d_L.p, h->d_U.p, h->d_D.p);
                else if (tma)
                    k_ldl_update_tma<64><<<dim3(gx, ny), tma_threads(64), tma_gemm_smem(64) + dbg_pad, st>>>(
                        ds, bl, h->d_tmaps64.p, h->d_tmap_of.p, mode, J0, na, h->tma_kmajor | dbg_fence, h->d_L.p, h->d_U.p, h->d_D.p);
                else
                    k_ldl_update_ldg<<<dim3(gx, ny), 256, 0, st>>>(ds, bl, mode, J0, na, h->d_L.p, h->d_U.p, h->d_D.p);
                LAUNCH(h);
            };
            for (int kb = 0; kb < B.maxns; kb += PB) {
                const int nact = B.active(kb);
                if (nact <= 0) break;
                { FineScope fs(h, Timers::F_DIAG);
                  k_piv_diag<<<nact, 256, 0, st>>>(ds, bl, kb, h->d_L.p, h->d_W.p, wo, h->d_D.p, h->d_Dinv.p, rp, h->d_nreg.p);
                  LAUNCH(h); }
                const int rows_below = B.maxnf - kb - 1;
                if (rows_below > 0) {
                    FineScope fs(h, Timers::F_ROWS);
                    k_piv_rows<<<dim3(nblk(rows_below, GBM), nact), 256, smrows, st>>>(ds, bl, kb, h->d_L.p, h->d_W.p, wo, h->d_Dinv.p);
                    LAUNCH(h);
                }
                if (kb + PB < B.maxns) { FineScope fs(h, Timers::F_PANELUPD); launch_update(0, kb); }
            }
            if (h->detail) { h->tm.end(st); h->tm.begin(Timers::SCHUR, st); }
            { FineScope fs(h, Timers::F_SCHUR); launch_update(1, 0); }
            if (h->detail) h->tm.end(st);
            FineScope fs(h, Timers::F_FINISH);
            k_finish_large<<<dim3(nblk(B.maxns, PB), B.cnt), 256, 0, st>>>(ds, bl, h->d_L.p, h->d_W.p, wo);
            LAUNCH(h);
        }
        if (ms) h->join(1);
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}

int factor(cb200_handle* h, bool static_reg) {
    h->tm.begin(Timers::FACTOR, h->stream);
    int rc = run_captured(h, h->g_factor[static_reg ? 1 : 0], [&]() { return factor_body(h, static_reg); }, 2);
    h->tm.end(h->stream);
    h->tm.nfactor += 1;
    return rc;
}

// y (permuted, in d_y) <- K^-1 : forward, diagonal, backward
SubTrees subtrees(cb200_handle* h) {
    SubTrees T;
    T.sn0 = h->d_sub_sn0.p; T.sn1 = h->d_sub_sn1.p; T.order = h->d_sub_order.p; T.order_ptr = h->d_sub_optr.p;
    T.lvl_off = h->d_sub_lvl.p; T.nlev = h->sub_nlev; T.maxcols = h->sub_maxcols; T.maxrows = h->sub_maxrows;
    return T;
}
size_t subtree_smem_fwd(const cb200_handle* h) { return ((size_t)h->sub_maxcols + h->sub_maxrows + 8 * SUB_MAXNF) * sizeof(double); }
size_t subtree_smem_bwd(const cb200_handle* h) { return ((size_t)h->sub_maxcols + 8 * (SUB_MAXNF + 32 * BT_LD)) * sizeof(double); }

int sweeps_body(cb200_handle* h) {
    const auto& S = h->S;
    if (h->nsub) {        // the bottom subtrees first: everything above them depends on their contributions
        FineScope fs(h, Timers::F_FWD_SUB);
        k_fwd_subtree<<<h->nsub, 256, subtree_smem_fwd(h), h->stream>>>(devsym(h), subtrees(h), h->d_L.p, h->d_y.p, h->d_uvec.p);
        LAUNCH(h);
    }
    for (int lv = 0; lv < S.nlevels; ++lv) launch_fwd_level(h, h->plan[lv], lv);
    for (int lv = S.nlevels - 1; lv >= 0; --lv) launch_bwd_level(h, h->plan[lv]);
    if (h->nsub) {
        FineScope fs(h, Timers::F_BWD_SUB);
        k_bwd_subtree<<<h->nsub, 256, subtree_smem_bwd(h), h->stream>>>(devsym(h), subtrees(h), h->d_L.p, h->d_Dinv.p, h->d_y.p);
        LAUNCH(h);
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}

int tri_solve(cb200_handle* h, const double* d_rhs, double* d_sol) {
    cudaStream_t st = h->stream;
    h->tm.begin(Timers::SOLVE, st);
    {
        FineScope fs(h, Timers::F_PERMUTE);
        if (h->N) { k_pack_perm<<<nblk(h->N, 256), 256, 0, st>>>(d_rhs, h->d_perm.p, h->N, h->d_y.p); LAUNCH(h); }
        if (h->dist && h->N) {
            // top-front entries of the right-hand side are contributed by rank 0 only
            k_mask_vec<<<nblk(h->N, 256), 256, 0, st>>>(h->d_y.p, h->d_topcolkeep.p, h->N); LAUNCH(h);
        }
    }
    int rc = run_captured(h, h->g_solve, [&]() { return sweeps_body(h); });
    if (rc) return rc;
    if (h->dist && h->N) {
        // every rank holds the solution on its own subtrees (+ the replicated top part): gather by
        // zeroing what a rank does not own and summing over NVLink
        FineScope fs(h, Timers::F_COMM);
        k_mask_vec<<<nblk(h->N, 256), 256, 0, st>>>(h->d_y.p, h->d_keepcol.p, h->N); LAUNCH(h);
        NCCL_OK(g_nccl.AllReduce(h->d_y.p, h->d_y.p, (size_t)h->N, ncclDouble, ncclSum, h->comm, st));
    }
    if (h->N) {
        FineScope fs(h, Timers::F_PERMUTE);
        k_unpack_perm<<<nblk(h->N, 256), 256, 0, st>>>(h->d_y.p, h->d_perm.p, h->N, d_sol); LAUNCH(h);
    }
    h->tm.end(st);
    h->tm.nsolve += 1;
    CUDA_OK(cudaGetLastError());
    return 0;
}

// e = b - K xi ; norm bits accumulated into d_scal[slot]
int residual(cb200_handle* h, const double* d_xi, double* d_e, int slot) {
    cudaStream_t st = h->stream;
    h->tm.begin(Timers::SPMV, st);
    CUDA_OK(cudaMemsetAsync(h->d_scal.p + slot, 0, sizeof(unsigned long long), st));
    if (h->N) {
        // lanes per row chosen from the mean row length of the symmetric matrix
        const double mean_row = 2.0 * (double)h->nnzK / (double)h->N;
        if (mean_row > 48.0)
            k_residual<32><<<nblk(h->N * 32, 256), 256, 0, st>>>(h->N, h->d_cp.p, h->d_ri.p, h->d_nz.p, h->d_tp.p,
                                                                 h->d_tc.p, h->d_tpos.p, d_xi, h->d_b.p, d_e,
                                                                 h->d_scal.p + slot);
        else
            k_residual<4><<<nblk(h->N * 4, 256), 256, 0, st>>>(h->N, h->d_cp.p, h->d_ri.p, h->d_nz.p, h->d_tp.p,
                                                               h->d_tc.p, h->d_tpos.p, d_xi, h->d_b.p, d_e,
                                                               h->d_scal.p + slot);
        if (h->n_long_rows) {
            k_residual_long<<<h->n_long_rows, 256, 0, st>>>(h->d_long_rows.p, h->d_cp.p, h->d_ri.p, h->d_nz.p,
                                                            h->d_tp.p, h->d_tc.p, h->d_tpos.p, d_xi, h->d_b.p,
                                                            d_e, h->d_scal.p + slot);
            LAUNCH(h);
        }
        LAUNCH(h);
    }
    h->tm.end(st);
    return 0;
}

int read_scalars(cb200_handle* h, double* out, int first, int count) {
    unsigned long long bits[4];
    CUDA_OK(cudaMemcpyAsync(bits, h->d_scal.p + first, count * sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < count; ++i) { long long b = (long long)bits[i]; std::memcpy(&out[i], &b, 8); }
    return 0;
}

}  // namespace

extern "C" {

int32_t cb200_create(int64_t N, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                     const int64_t* Dsigns, const cb200_settings* stp, cb200_handle** out) {
    try {
        cb200_settings st;
        if (stp) st = *stp; else cb200_default_settings(&st);
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            set_error("cb200_create: no CUDA device available (this backend has no CPU fallback)");
            return -3;
        }
        CUDA_OK(cudaSetDevice(st.device));
        auto* h = new cb200_handle();
        struct Guard { cb200_handle* h; ~Guard() { if (h) cb200_destroy(h); } } guard{h};   // error paths free everything
        h->st = st; h->N = N; h->base = st.index_base;
        const int64_t base = st.index_base;
        std::vector<int64_t> cp(N + 1), ri;
        for (int64_t i = 0; i <= N; ++i) cp[i] = colptr[i] - base;
        const int64_t nnz = cp[N];
        h->nnzK = nnz;
        ri.resize(nnz);
        for (int64_t i = 0; i < nnz; ++i) ri[i] = rowval[i] - base;
        if (!g_block_hint.empty() && (int64_t)g_block_hint.size() != N) g_block_hint.clear();
        h->opt = options_from_settings(&st);
        // ordering = 1 (auto) with dense cone blocks (PSD cones) in K: the blocks are detected here from the
        // pattern and Dsigns alone - a maximal run of columns of the (-,-) part in which column j holds ALL
        // earlier columns of the run as rows (a dense triangle; dense SOC blocks have at most 4 columns) - so
        // that every caller, the Julia shim included, gets the PSD-safe ordering without knowing about it:
        // coupled variables first, then the blocks by nested dissection of the block graph, or the
        // AMD-class order when there are only a few blocks (symbolic.cpp / ordering.cpp).
        std::vector<int32_t> detected;
        if (h->opt.ordering == 1 && Dsigns && !h->opt.block_id) {
            std::vector<int32_t> start(N, -1);                 // start[j] = first column of the run j belongs to
            for (int64_t j = 0; j < N; ++j) {
                if (Dsigns[j] >= 0) continue;
                start[j] = (int32_t)j;
                if (j == 0 || start[j - 1] < 0) continue;
                const int64_t want = j - start[j - 1];         // rows start[j-1] .. j-1 must all be present
                int64_t run = 0;
                for (int64_t p = cp[j + 1] - 1; p >= cp[j]; --p) {
                    const int64_t r = ri[p];
                    if (r == j) continue;
                    if (r == j - 1 - run) ++run; else if (r < j - 1 - run) break;
                    if (run == want) break;
                }
                if (run == want) start[j] = start[j - 1];
            }
            detected.assign(N, -1);
            int32_t nblk_found = 0;
            for (int64_t j = 0; j < N;) {
                if (start[j] < 0) { ++j; continue; }
                int64_t e = j; while (e + 1 < N && start[e + 1] == start[j]) ++e;
                if (e - j + 1 >= 5) { for (int64_t q = j; q <= e; ++q) detected[q] = nblk_found; ++nblk_found; }
                j = e + 1;
            }
            if (nblk_found > 0) h->opt.block_id = detected.data();
        }
        symbolic_analyze(N, cp.data(), ri.data(), h->opt, nullptr, h->S);
        h->opt.block_id = nullptr;
        g_block_hint.clear();
        const Symbolic& S = h->S;
        CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        for (int i = 0; i < 3; ++i) {
            CUDA_OK(cudaStreamCreateWithFlags(&h->side[i], cudaStreamNonBlocking));
            CUDA_OK(cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming));
        }
        CUDA_OK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        if (const char* e = getenv("CB200_MULTISTREAM")) h->multi_stream = e[0] != '0';
        if (const char* e = getenv("CB200_SORT_BATCHES")) h->sort_batches = e[0] != '0';
        if (const char* e = getenv("CB200_SUBTREE")) h->use_subtree = e[0] == '1';
        if (const char* e = getenv("CB200_MERGED_ROWS")) h->merged_rows = e[0] != '0';
        cudaStream_t s = h->stream;
        // ---- K
        std::vector<int32_t> ri32(ri.begin(), ri.end());
        CUDA_OK(h->d_cp.upload(cp, s)); CUDA_OK(h->d_ri.upload(ri32, s));
        CUDA_OK(h->d_nz.alloc(nnz));
        if (nnz) CUDA_OK(cudaMemcpyAsync(h->d_nz.p, nzval, nnz * sizeof(double), cudaMemcpyHostToDevice, s));
        // row-wise index of the strictly-upper entries + diagonal positions
        std::vector<int64_t> tp(N + 1, 0), diagidx(N, -1);
        for (int64_t j = 0; j < N; ++j)
            for (int64_t p = cp[j]; p < cp[j + 1]; ++p) {
                if (ri[p] == j) diagidx[j] = p; else if (ri[p] < j) tp[ri[p] + 1]++;
            }
        for (int64_t j = 0; j < N; ++j) tp[j + 1] += tp[j];
        std::vector<int32_t> tc(tp[N]); std::vector<int64_t> tpos(tp[N]);
        {
            std::vector<int64_t> pos(tp.begin(), tp.end() - 1);
            for (int64_t j = 0; j < N; ++j)
                for (int64_t p = cp[j]; p < cp[j + 1]; ++p)
                    if (ri[p] < j) { int64_t q = pos[ri[p]]++; tc[q] = (int32_t)j; tpos[q] = p; }
        }
        CUDA_OK(h->d_tp.upload(tp, s)); CUDA_OK(h->d_tc.upload(tc, s)); CUDA_OK(h->d_tpos.upload(tpos, s));
        {
            std::vector<int32_t> longrows;
            for (int64_t j = 0; j < N; ++j)
                if ((cp[j + 1] - cp[j]) + (tp[j + 1] - tp[j]) > LONG_ROW) longrows.push_back((int32_t)j);
            h->n_long_rows = (int32_t)longrows.size();
            CUDA_OK(h->d_long_rows.upload(longrows, s));
        }
        h->have_diag = std::all_of(diagidx.begin(), diagidx.end(), [](int64_t v) { return v >= 0; });
        if (h->have_diag) CUDA_OK(h->d_diagidx.upload(diagidx, s));
        CUDA_OK(h->d_amap.upload(S.a_map, s));
        std::vector<int8_t> dsp(N), dso(N);
        for (int64_t k = 0; k < N; ++k) {
            dso[k] = (int8_t)(Dsigns ? (Dsigns[k] >= 0 ? 1 : -1) : 1);
        }
        for (int64_t k = 0; k < N; ++k) dsp[k] = dso[S.perm[k]];
        CUDA_OK(h->d_dsign_perm.upload(dsp, s)); CUDA_OK(h->d_dsign_orig.upload(dso, s));
        CUDA_OK(h->d_perm.upload(S.perm, s));
        // ---- symbolic
        CUDA_OK(h->d_sn_first.upload(S.sn_first, s)); CUDA_OK(h->d_rows_ptr.upload(S.rows_ptr, s));
        CUDA_OK(h->d_rows.upload(S.rows, s)); CUDA_OK(h->d_rel.upload(S.rel, s));
        CUDA_OK(h->d_child_ptr.upload(S.child_ptr, s)); CUDA_OK(h->d_child_list.upload(S.child_list, s));
        CUDA_OK(h->d_panel_off.upload(S.panel_off, s)); CUDA_OK(h->d_upd_off.upload(S.upd_off, s));
        CUDA_OK(h->d_ld.upload(S.panel_ld, s));
        CUDA_OK(h->d_front_ptr.upload(S.front_ptr, s)); CUDA_OK(h->d_asm_base.upload(S.asm_base, s));
        CUDA_OK(h->d_asm_colptr.upload(S.asm_colptr, s)); CUDA_OK(h->d_asm_src.upload(S.asm_src, s));
        CUDA_OK(h->d_asm_child.upload(S.asm_child, s));
        // ---- level plans (+ workspace sized for them)
        { int rcp = build_plans(h); if (rcp) return rcp; }
        // ---- numeric storage
        CUDA_OK(h->d_L.alloc((size_t)S.panel_off.back()));
        CUDA_OK(h->d_U.alloc((size_t)std::max<int64_t>(1, S.upd_total)));
        CUDA_OK(h->d_D.alloc(N)); CUDA_OK(h->d_Dinv.alloc(N));
        if (N) { CUDA_OK(cudaMemsetAsync(h->d_D.p, 0, N * sizeof(double), s)); CUDA_OK(cudaMemsetAsync(h->d_Dinv.p, 0, N * sizeof(double), s)); }
        CUDA_OK(h->d_uvec.alloc(std::max<size_t>(1, S.rows.size())));

        for (DevBuf<double>* b : {&h->d_b, &h->d_x, &h->d_e, &h->d_dx, &h->d_y}) {
            CUDA_OK(b->alloc(std::max<int64_t>(1, N)));
            CUDA_OK(cudaMemsetAsync(b->p, 0, std::max<int64_t>(1, N) * sizeof(double), s));
        }
        CUDA_OK(h->d_eps.alloc(1)); CUDA_OK(h->d_scal.alloc(4)); CUDA_OK(h->d_nreg.alloc(1)); CUDA_OK(h->d_reglog.alloc(64));
        CUDA_OK(cudaMemsetAsync(h->d_eps.p, 0, sizeof(double), s));
        CUDA_OK(cudaMemsetAsync(h->d_nreg.p, 0, sizeof(unsigned int), s));
        // opt in to large dynamic shared memory for the bigger small-front classes
        CUDA_OK(cudaFuncSetAttribute(k_factor_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        CUDA_OK(cudaFuncSetAttribute(k_factor_small<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (152 * 152 + SB * SB) * (int)sizeof(double)));
        CUDA_OK(cudaFuncSetAttribute(k_factor_small<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (64 * 64 + SB * SB) * (int)sizeof(double)));
        CUDA_OK(cudaFuncSetAttribute(k_factor_small<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (32 * 32 + SB * SB) * (int)sizeof(double)));
        CUDA_OK(cudaFuncSetAttribute(k_piv_rows, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (2 * PB * (PB + 1)) * (int)sizeof(double)));
        CUDA_OK(cudaFuncSetAttribute(k_ldl_update_tma<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_gemm_smem(128)));
        {   // (CB200_TMA_ATTR_KB: debugging aid, see DESIGN.md section 8 "reproducibility")
            const char* e = getenv("CB200_TMA_ATTR_KB");
            const int attr = e ? atoi(e) * 1024 + 64 : 200 * 1024;
            CUDA_OK(cudaFuncSetAttribute(k_ldl_update_tma<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, attr));
        }
        { int rct = build_tensor_maps(h); if (rct) return rct; }
        CUDA_OK(cudaFuncSetAttribute(k_fwd_subtree, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        CUDA_OK(cudaFuncSetAttribute(k_bwd_subtree, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        CUDA_OK(cudaFuncSetAttribute(k_bwd_warp<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CUDA_OK(cudaFuncSetAttribute(k_bwd_warp<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CUDA_OK(cudaFuncSetAttribute(k_bwd_warp<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CUDA_OK(cudaFuncSetAttribute(k_fwd_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CUDA_OK(cudaFuncSetAttribute(k_bwd_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (S.max_front > 25000) { set_error("front too large for the single-CTA solve kernels"); return -4; }
        CUDA_OK(cudaStreamSynchronize(s));
        guard.h = nullptr;
        *out = h;
        return 0;
    } catch (const std::exception& e) { set_error(e.what()); return -1; }
}

void cb200_destroy(cb200_handle* h) {
    if (!h) return;
    if (h->stream) { cudaStreamSynchronize(h->stream); }
    // graphs captured with NCCL kernels inside hold references on the communicator: ncclCommDestroy waits
    // for them, so the executables go first (seen as a teardown hang with CB200_DIST_GRAPH=1)
    for (GraphExec* g : {&h->g_factor[0], &h->g_factor[1], &h->g_solve})
        if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; }
    if (h->comm && g_nccl.ok) { g_nccl.CommDestroy(h->comm); h->comm = nullptr; }
    cudaStream_t s = h->stream;
    for (int i = 0; i < 3; ++i) {
        if (h->side[i]) { cudaStreamSynchronize(h->side[i]); cudaStreamDestroy(h->side[i]); }
        if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    delete h;
    if (s) cudaStreamDestroy(s);
}

// Staging for the inner-boundary uploads.  The reference issues ~5 update_values!/scale_values! calls
// per sparse cone per iteration (directldl_datamaps.jl:61-79), so these calls must not allocate or
// synchronise: the device staging buffers only grow, and the stream is ordered, so the kernel of
// call k+1 cannot overtake the copy of call k.  The staging buffer is reused by the next call only
// after the previous kernel was enqueued behind its copy on the same stream; the H2D copies from
// pageable host memory return after the source has been staged by the driver, so the caller may
// overwrite `index` / `values` on return.
// Page-locked caller buffers make cudaMemcpyAsync truly asynchronous: then (and only then) the call
// must wait for the copy before returning, because the caller may overwrite its buffer.
static bool host_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

static int stage_reserve(cb200_handle* h, size_t nidx, size_t nval) {
    if (nidx > h->stage_idx_cap) {
        CUDA_OK(cudaStreamSynchronize(h->stream));          // the old buffer may still be in use
        size_t cap = std::max(nidx, std::max<size_t>(4096, 2 * h->stage_idx_cap));
        CUDA_OK(h->d_stage_idx.alloc(cap)); h->stage_idx_cap = cap;
    }
    if (nval > h->stage_val_cap) {
        CUDA_OK(cudaStreamSynchronize(h->stream));
        size_t cap = std::max(nval, std::max<size_t>(4096, 2 * h->stage_val_cap));
        CUDA_OK(h->d_stage_val.alloc(cap)); h->stage_val_cap = cap;
    }
    return 0;
}

int32_t cb200_update_values(cb200_handle* h, const int64_t* index, const double* values, int64_t len) {
    if (len <= 0) return 0;
    CUDA_OK(cudaSetDevice(h->st.device));
    { int rc = stage_reserve(h, (size_t)len, (size_t)len); if (rc) return rc; }
    CUDA_OK(cudaMemcpyAsync(h->d_stage_idx.p, index, len * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaMemcpyAsync(h->d_stage_val.p, values, len * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    k_update_values<<<nblk(len, 256), 256, 0, h->stream>>>(h->d_nz.p, h->d_stage_idx.p, h->d_stage_val.p, len, h->base);
    LAUNCH(h);
    CUDA_OK(cudaGetLastError());
    if (host_is_pinned(index) || host_is_pinned(values)) CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}

int32_t cb200_scale_values(cb200_handle* h, const int64_t* index, int64_t len, double scale) {
    if (len <= 0) return 0;
    CUDA_OK(cudaSetDevice(h->st.device));
    { int rc = stage_reserve(h, (size_t)len, 0); if (rc) return rc; }
    CUDA_OK(cudaMemcpyAsync(h->d_stage_idx.p, index, len * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    k_scale_values<<<nblk(len, 256), 256, 0, h->stream>>>(h->d_nz.p, h->d_stage_idx.p, scale, len, h->base);
    LAUNCH(h);
    CUDA_OK(cudaGetLastError());
    if (host_is_pinned(index)) CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}

static int finish_factor(cb200_handle* h) {
    // success = all(isfinite, Dinv)   (directldl_qdldl.jl:79)
    CUDA_OK(cudaMemsetAsync(h->d_scal.p + 3, 0, sizeof(unsigned long long), h->stream));
    if (h->N) { k_absmax<<<std::min(nblk(h->N, 256), 1184), 256, 0, h->stream>>>(h->d_Dinv.p, h->N, h->d_scal.p + 3); LAUNCH(h); }
    if (h->dist) {
        // every rank must take the same success / failure branch: reduce the status over the ranks
        // (the value is the bit pattern of a non-negative double, so unsigned max == double max)
        NCCL_OK(g_nccl.AllReduce(h->d_scal.p + 3, h->d_scal.p + 3, 1, ncclUint64, ncclMax, h->comm, h->stream));
    }
    double v[1];
    int rc = read_scalars(h, v, 3, 1);
    if (rc) return rc;
    h->tm.collect();
    return std::isfinite(v[0]) ? 0 : 1;
}

int32_t cb200_refactor(cb200_handle* h) {
    CUDA_OK(cudaSetDevice(h->st.device));
    int rc = factor(h, /*static_reg=*/false);
    if (rc) return rc;
    return finish_factor(h);
}

int32_t cb200_solve(cb200_handle* h, double* x, const double* b) {
    const int64_t N = h->N;
    if (N == 0) return 0;
    CUDA_OK(cudaSetDevice(h->st.device));
    CUDA_OK(cudaMemcpyAsync(h->d_b.p, b, N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    int rc = tri_solve(h, h->d_b.p, h->d_x.p);
    if (rc) return rc;
    CUDA_OK(cudaMemcpyAsync(x, h->d_x.p, N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->tm.collect();
    return 0;
}

int32_t cb200_info(const cb200_handle* h, int64_t* nnzA, int64_t* nnzL, int32_t* ngpus) {
    if (nnzA) *nnzA = h->nnzK;
    if (nnzL) *nnzL = h->S.nnzL;
    if (ngpus) *ngpus = h->nranks;
    return 0;
}

int32_t cb200_set_maps(cb200_handle* h, int64_t n, int64_t m, int64_t p,
                       const int64_t* map_P, int64_t nnzP, const int64_t* map_A, int64_t nnzA,
                       const int64_t* map_Hs, int64_t nHs, const int64_t* map_diag_full,
                       int64_t ncones, const int32_t* cone_type, const int64_t* cone_dim,
                       const int64_t* map_soc_u, const int64_t* map_soc_v, const int64_t* map_soc_D) {
    try {
        if (n + m + p != h->N) { set_error("cb200_set_maps: n+m+p != N"); return -2; }
        CUDA_OK(cudaSetDevice(h->st.device));
        const int64_t base = h->base;
        cudaStream_t s = h->stream;
        h->n = n; h->m = m; h->p = p;
        auto rebased = [&](const int64_t* src, int64_t len) {
            std::vector<int64_t> v(len);
            for (int64_t i = 0; i < len; ++i) v[i] = src[i] - base;
            return v;
        };
        CUDA_OK(h->d_mapP.upload(rebased(map_P, nnzP), s));
        CUDA_OK(h->d_mapA.upload(rebased(map_A, nnzA), s));
        std::vector<int64_t> mapHs = rebased(map_Hs, nHs);
        CUDA_OK(h->d_mapHs.upload(mapHs, s));
        CUDA_OK(h->d_diagidx.upload(rebased(map_diag_full, h->N), s));
        h->have_diag = true;
        // cone tables
        std::vector<int8_t> dg_kind; std::vector<int32_t> dg_midx, dg_cone; std::vector<int64_t> dg_map;
        std::vector<int32_t> sd_moff, sd_dim, sd_socid; std::vector<int64_t> sd_hoff;
        std::vector<int32_t> ex_src, ex_cone, exD_cone;
        std::vector<int32_t> psd_side; std::vector<int64_t> psd_roff, psd_hoff;
        int64_t moff = 0, hoff = 0, socrow = 0, roff = 0;
        int32_t socid = 0;
        h->psd_maxn = 0;
        for (int64_t i = 0; i < ncones; ++i) {
            const int t = cone_type[i];
            const int64_t d = cone_dim[i];
            if (t == 0 || t == 1) {
                for (int64_t k = 0; k < d; ++k) {
                    dg_kind.push_back((int8_t)t); dg_midx.push_back((int32_t)(moff + k));
                    dg_cone.push_back(0); dg_map.push_back(mapHs[hoff + k]);
                }
                moff += d; hoff += d;
            } else if (t == 2) {
                if (d > 4) {
                    for (int64_t k = 0; k < d; ++k) {
                        dg_kind.push_back(k == 0 ? 3 : 2); dg_midx.push_back((int32_t)(moff + k));
                        dg_cone.push_back(socid); dg_map.push_back(mapHs[hoff + k]);
                        ex_src.push_back((int32_t)(socrow + k)); ex_cone.push_back(socid);
                    }
                    exD_cone.push_back(socid);
                    hoff += d;
                } else {
                    sd_moff.push_back((int32_t)moff); sd_dim.push_back((int32_t)d);
                    sd_socid.push_back(socid); sd_hoff.push_back(hoff);
                    hoff += d * (d + 1) / 2;
                }
                moff += d; socrow += d; socid++;
            } else if (t == 3) {
                const int64_t ne = d * (d + 1) / 2;
                psd_side.push_back((int32_t)d); psd_roff.push_back(roff); psd_hoff.push_back(hoff);
                roff += d * d; hoff += ne * (ne + 1) / 2; moff += ne;
                h->psd_maxn = std::max<int32_t>(h->psd_maxn, (int32_t)d);
            } else if (t == 4 || t == 5) {
                // exponential / power cone: dense 3x3 block whose values the caller computes
                // (coneops_expcone.jl:92-100) and sends through cb200_update_values before
                // cb200_update_cones; no cone kernel touches these entries
                if (d != 3) { set_error("cb200_set_maps: exponential / power cones have dimension 3"); return -2; }
                moff += 3; hoff += 6;
            } else if (t == 6) {
                // generalised power cone: diagonal block + expansion columns, all sent by the
                // caller through cb200_update_values (coneops_genpowcone.jl:91-108,
                // directldl_datamaps.jl:146-166)
                moff += d; hoff += d;
            } else { set_error("cb200_set_maps: unsupported cone type"); return -2; }
        }
        if (moff != m || hoff != nHs) { set_error("cb200_set_maps: cone table inconsistent with m / nHs"); return -2; }
        h->ndiag = (int64_t)dg_kind.size();
        CUDA_OK(h->d_dg_kind.upload(dg_kind, s)); CUDA_OK(h->d_dg_midx.upload(dg_midx, s));
        CUDA_OK(h->d_dg_cone.upload(dg_cone, s)); CUDA_OK(h->d_dg_map.upload(dg_map, s));
        h->nsocd = (int32_t)sd_dim.size();
        CUDA_OK(h->d_sd_moff.upload(sd_moff, s)); CUDA_OK(h->d_sd_dim.upload(sd_dim, s));
        CUDA_OK(h->d_sd_socid.upload(sd_socid, s)); CUDA_OK(h->d_sd_hoff.upload(sd_hoff, s));
        h->nexp = (int64_t)ex_src.size(); h->nsocs = (int32_t)exD_cone.size();
        CUDA_OK(h->d_ex_src.upload(ex_src, s)); CUDA_OK(h->d_ex_cone.upload(ex_cone, s));
        CUDA_OK(h->d_exD_cone.upload(exD_cone, s));
        CUDA_OK(h->d_mapu.upload(rebased(map_soc_u, h->nexp), s));
        CUDA_OK(h->d_mapv.upload(rebased(map_soc_v, h->nexp), s));
        CUDA_OK(h->d_mapD.upload(rebased(map_soc_D, 2 * (int64_t)h->nsocs), s));
        h->npsd = (int32_t)psd_side.size(); h->psd_rtotal = roff;
        CUDA_OK(h->d_psd_side.upload(psd_side, s)); CUDA_OK(h->d_psd_roff.upload(psd_roff, s));
        CUDA_OK(h->d_psd_hoff.upload(psd_hoff, s));
        CUDA_OK(h->d_psd_R.alloc(std::max<int64_t>(1, roff))); CUDA_OK(h->d_psd_A.alloc(std::max<int64_t>(1, roff)));
        h->nsoc = socid; h->nsocrows = socrow;
        CUDA_OK(h->d_w.alloc(std::max<int64_t>(1, m)));
        CUDA_OK(h->d_eta.alloc(std::max<int64_t>(1, h->nsoc))); CUDA_OK(h->d_socd.alloc(std::max<int64_t>(1, h->nsoc)));
        CUDA_OK(h->d_socu.alloc(std::max<int64_t>(1, socrow))); CUDA_OK(h->d_socv.alloc(std::max<int64_t>(1, socrow)));
        CUDA_OK(h->d_rx.alloc(std::max<int64_t>(1, n))); CUDA_OK(h->d_rz.alloc(std::max<int64_t>(1, m)));
        if (h->psd_maxn) CUDA_OK(cudaFuncSetAttribute(k_psd_skron, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      std::max(48 * 1024, h->psd_maxn * h->psd_maxn * 8)));
        CUDA_OK(cudaStreamSynchronize(s));
        h->maps_set = true;
        return 0;
    } catch (const std::exception& e) { set_error(e.what()); return -1; }
}

int32_t cb200_update_cones(cb200_handle* h, const double* w, const double* soc_eta,
                           const double* soc_d, const double* soc_u, const double* soc_v,
                           const double* psd_R) {
    if (!h->maps_set) { set_error("cb200_update_cones: cb200_set_maps not called"); return -2; }
    CUDA_OK(cudaSetDevice(h->st.device));
    cudaStream_t st = h->stream;
    h->tm.begin(Timers::CONE, st);
    // host pointers normally; in resident mode (cb200_set_resident) non-NULL arguments are DEVICE pointers
    // (inputs staged in HBM by the caller, copied device-to-device) and NULL keeps the previous state
    const cudaMemcpyKind kind = h->resident ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    auto put = [&](double* dst, const double* src, int64_t len) -> cudaError_t {
        if (len <= 0 || (h->resident && !src)) return cudaSuccess;
        return cudaMemcpyAsync(dst, src, len * sizeof(double), kind, st);
    };
    CUDA_OK(put(h->d_w.p, w, h->m));
    CUDA_OK(put(h->d_eta.p, soc_eta, h->nsoc)); CUDA_OK(put(h->d_socd.p, soc_d, h->nsoc));
    CUDA_OK(put(h->d_socu.p, soc_u, h->nsocrows)); CUDA_OK(put(h->d_socv.p, soc_v, h->nsocrows));
    CUDA_OK(put(h->d_psd_R.p, psd_R, h->psd_rtotal));
    if (h->ndiag) {
        k_hs_diag<<<nblk(h->ndiag, 256), 256, 0, st>>>(h->ndiag, h->d_dg_kind.p, h->d_dg_midx.p, h->d_dg_cone.p,
                                                       h->d_dg_map.p, h->d_w.p, h->d_eta.p, h->d_socd.p, h->d_nz.p);
        LAUNCH(h);
    }
    if (h->nsocd) {
        k_hs_soc_dense<<<nblk(h->nsocd, 128), 128, 0, st>>>(h->nsocd, h->d_sd_moff.p, h->d_sd_dim.p, h->d_sd_socid.p,
                                                            h->d_sd_hoff.p, h->d_mapHs.p, h->d_w.p, h->d_eta.p, h->d_nz.p);
        LAUNCH(h);
    }
    if (h->nexp) {
        k_soc_expansion<<<nblk(h->nexp, 256), 256, 0, st>>>(h->nexp, h->d_ex_src.p, h->d_ex_cone.p, h->d_mapu.p,
                                                            h->d_mapv.p, h->d_socu.p, h->d_socv.p, h->d_eta.p, h->d_nz.p);
        k_soc_D<<<nblk(h->nsocs, 128), 128, 0, st>>>(h->nsocs, h->d_exD_cone.p, h->d_mapD.p, h->d_eta.p, h->d_nz.p);
        h->tm.nlaunch += 2;
    }
    if (h->npsd) {
        k_psd_rrt<<<h->npsd, 256, 0, st>>>(h->d_psd_side.p, h->d_psd_roff.p, h->d_psd_R.p, h->d_psd_A.p);
        const int ne = h->psd_maxn * (h->psd_maxn + 1) / 2;
        k_psd_skron<<<dim3(std::max(1, std::min(64, nblk(ne, 8))), h->npsd), 256,
                      (size_t)h->psd_maxn * h->psd_maxn * sizeof(double), st>>>(
            h->d_psd_side.p, h->d_psd_roff.p, h->d_psd_hoff.p, h->d_psd_A.p, h->d_mapHs.p, h->d_nz.p);
        h->tm.nlaunch += 2;
    }
    h->tm.end(st);
    CUDA_OK(cudaGetLastError());
    int rc = factor(h, /*static_reg=*/true);
    if (rc) return rc;
    return finish_factor(h);
}

int32_t cb200_setrhs(cb200_handle* h, const double* rhsx, const double* rhsz) {
    if (!h->maps_set) { set_error("cb200_setrhs: cb200_set_maps not called"); return -2; }
    CUDA_OK(cudaSetDevice(h->st.device));
    cudaStream_t st = h->stream;
    // resident mode: non-NULL arguments are device pointers (see cb200_update_cones), stream-ordered copy
    const cudaMemcpyKind kind = h->resident ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (h->n && rhsx) CUDA_OK(cudaMemcpyAsync(h->d_rx.p, rhsx, h->n * sizeof(double), kind, st));
    if (h->m && rhsz) CUDA_OK(cudaMemcpyAsync(h->d_rz.p, rhsz, h->m * sizeof(double), kind, st));
    if (!h->resident) CUDA_OK(cudaStreamSynchronize(st));        // the caller may overwrite its buffers on return
    return 0;
}

int32_t cb200_solve_ir(cb200_handle* h, const double* rhsx, const double* rhsz,
                       double* lhsx, double* lhsz, int32_t* ir_rounds) {
    if (!h->maps_set) { set_error("cb200_solve_ir: cb200_set_maps not called"); return -2; }
    CUDA_OK(cudaSetDevice(h->st.device));
    cudaStream_t st = h->stream;
    const int64_t N = h->N, n = h->n, m = h->m;
    if (ir_rounds) *ir_rounds = 0;
    if (N == 0) return 0;
    if (!h->resident) {      // each part is optional on its own (a NULL part keeps what cb200_setrhs stored)
        if (n && rhsx) CUDA_OK(cudaMemcpyAsync(h->d_rx.p, rhsx, n * sizeof(double), cudaMemcpyHostToDevice, st));
        if (m && rhsz) CUDA_OK(cudaMemcpyAsync(h->d_rz.p, rhsz, m * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    k_build_rhs<<<nblk(N, 256), 256, 0, st>>>(h->d_rx.p, h->d_rz.p, n, m, N, h->d_b.p);
    CUDA_OK(cudaMemsetAsync(h->d_scal.p + 1, 0, sizeof(unsigned long long), st));
    k_absmax<<<std::min(nblk(N, 256), 1184), 256, 0, st>>>(h->d_b.p, N, h->d_scal.p + 1);
    h->tm.nlaunch += 2;
    double* x = h->d_x.p; double* dx = h->d_dx.p; double* e = h->d_e.p;
    int rc = tri_solve(h, h->d_b.p, x);
    if (rc) return rc;
    int ok = 1;
    if (h->st.iterative_refinement_enable) {
        // _iterative_refinement (kktsolver_directldl.jl:389-449)
        rc = residual(h, x, e, 2); if (rc) return rc;
        double sc[2];
        rc = read_scalars(h, sc, 1, 2); if (rc) return rc;
        const double normb = sc[0];
        double norme = sc[1];
        if (!std::isfinite(norme)) ok = 0;
        for (int i = 0; ok && i < h->st.iterative_refinement_max_iter; ++i) {
            if (norme <= h->st.iterative_refinement_abstol + h->st.iterative_refinement_reltol * normb) break;
            const double lastnorme = norme;
            rc = tri_solve(h, e, dx); if (rc) return rc;
            k_axpy1<<<nblk(N, 256), 256, 0, st>>>(dx, x, N); LAUNCH(h);
            rc = residual(h, dx, e, 2); if (rc) return rc;
            rc = read_scalars(h, sc, 2, 1); if (rc) return rc;
            norme = sc[0];
            if (ir_rounds) (*ir_rounds)++;
            if (!std::isfinite(norme)) { ok = 0; break; }
            const double ratio = lastnorme / norme;
            if (ratio < h->st.iterative_refinement_stop_ratio) {
                if (ratio > 1.0) std::swap(x, dx);
                break;
            }
            std::swap(x, dx);
        }
    } else {
        CUDA_OK(cudaMemsetAsync(h->d_scal.p + 2, 0, sizeof(unsigned long long), st));
        k_absmax<<<std::min(nblk(N, 256), 1184), 256, 0, st>>>(x, N, h->d_scal.p + 2); LAUNCH(h);
        double sc[1];
        rc = read_scalars(h, sc, 2, 1); if (rc) return rc;
        ok = std::isfinite(sc[0]);
    }
    // keep handle buffers consistent with the pointer swaps (kktsolver_directldl.jl:445)
    if (x != h->d_x.p) std::swap(h->d_x.p, h->d_dx.p);
    if (ok && !h->resident) {
        if (lhsx && n) CUDA_OK(cudaMemcpyAsync(lhsx, h->d_x.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (lhsz && m) CUDA_OK(cudaMemcpyAsync(lhsz, h->d_x.p + n, m * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    CUDA_OK(cudaStreamSynchronize(st));
    h->tm.collect();
    if (!ok && !h->resident) {
        // a failed solve must not leave stale finite values in the caller's vectors: the reference
        // checks all(isfinite, x) on them when refinement is disabled (kktsolver_directldl.jl:356-370)
        const double qnan = std::numeric_limits<double>::quiet_NaN();
        if (lhsx) std::fill(lhsx, lhsx + n, qnan);
        if (lhsz) std::fill(lhsz, lhsz + m, qnan);
    }
    return ok ? 0 : 1;
}

static int32_t update_block(cb200_handle* h, const DevBuf<int64_t>& map, const double* values, int64_t len) {
    if (!h->maps_set) { set_error("cb200_update_P/A: cb200_set_maps not called"); return -2; }
    if ((int64_t)map.n != len) { set_error("cb200_update_P/A: length mismatch"); return -2; }
    if (len == 0) return 0;
    CUDA_OK(cudaSetDevice(h->st.device));
    { int rc = stage_reserve(h, 0, (size_t)len); if (rc) return rc; }
    CUDA_OK(cudaMemcpyAsync(h->d_stage_val.p, values, len * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    k_update_values<<<nblk(len, 256), 256, 0, h->stream>>>(h->d_nz.p, map.p, h->d_stage_val.p, len, 0);
    LAUNCH(h);
    CUDA_OK(cudaGetLastError());
    if (host_is_pinned(values)) CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}
int32_t cb200_update_P(cb200_handle* h, const double* values, int64_t len) { return update_block(h, h->d_mapP, values, len); }
int32_t cb200_update_A(cb200_handle* h, const double* values, int64_t len) { return update_block(h, h->d_mapA, values, len); }

int32_t cb200_download(cb200_handle* h, int32_t what, double* out, int64_t len) {
    CUDA_OK(cudaSetDevice(h->st.device));
    cudaStream_t st = h->stream;
    const double* src = nullptr; int64_t n = 0;
    switch (what) {
        case 0: src = h->d_nz.p; n = h->nnzK; break;
        case 1: src = h->d_D.p; n = h->N; break;
        case 2: src = h->d_L.p; n = (int64_t)h->d_L.n; break;
        case 3: { if (len != h->N) { set_error("download: length"); return -2; }
                  for (int64_t i = 0; i < len; ++i) out[i] = (double)h->S.perm[i]; return 0; }
        case 4: src = h->d_eps.p; n = 1; break;
        case 5: { if (len != 1) { set_error("download: length"); return -2; }
                  unsigned int v = 0;
                  CUDA_OK(cudaMemcpyAsync(&v, h->d_nreg.p, sizeof(v), cudaMemcpyDeviceToHost, st));
                  CUDA_OK(cudaStreamSynchronize(st)); out[0] = (double)v; return 0; }
        case 6: src = h->d_x.p; n = h->N; break;       // full solution [x; z; expansion variables] of the last solve
        case 7: { if (len != 64) { set_error("download: length"); return -2; }     // ORIGINAL indices of the first 64 regularised pivots
                  int32_t v[64]; unsigned int cnt = 0;
                  CUDA_OK(cudaMemcpyAsync(v, h->d_reglog.p, sizeof(v), cudaMemcpyDeviceToHost, st));
                  CUDA_OK(cudaMemcpyAsync(&cnt, h->d_nreg.p, sizeof(cnt), cudaMemcpyDeviceToHost, st));
                  CUDA_OK(cudaStreamSynchronize(st));
                  for (int i = 0; i < 64; ++i) out[i] = i < (int)cnt ? (double)h->S.perm[v[i]] : -1.0;
                  return 0; }
        default: set_error("download: bad selector"); return -2;
    }
    if (len != n) { set_error("download: length mismatch"); return -2; }
    if (n) CUDA_OK(cudaMemcpyAsync(out, src, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int32_t cb200_get_timers(cb200_handle* h, double* out, int32_t len) {
    double v[11] = {h->tm.ms[0], h->tm.ms[1], h->tm.ms[2], h->tm.ms[3], h->tm.nfactor, h->tm.nsolve, h->tm.nlaunch,
                    h->tm.ms[4], h->tm.ms[5], h->tm.ms[6], h->tm.ms[7]};
    for (int i = 0; i < len && i < 11; ++i) out[i] = v[i];
    return 0;
}
int32_t cb200_nccl_unique_id(char out[128]) {
    if (!g_nccl.load()) { set_error("NCCL library not found (dlopen libnccl.so.2)"); return -5; }
    ncclUniqueId id;
    NCCL_OK(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId size");
    std::memcpy(out, &id, 128);
    return 0;
}

int32_t cb200_dist_init(cb200_handle* h, int32_t rank, int32_t nranks, const char uid[128]) {
    try {
        if (nranks <= 1) return 0;
        if (!g_nccl.load()) { set_error("NCCL library not found (dlopen libnccl.so.2)"); return -5; }
        CUDA_OK(cudaSetDevice(h->st.device));
        ncclUniqueId id; std::memcpy(&id, uid, 128);
        NCCL_OK(g_nccl.CommInitRank(&h->comm, nranks, id, rank));
        h->rank = rank; h->nranks = nranks; h->dist = true;
        const Symbolic& S = h->S;
        partition_subtrees(S, nranks, h->owner, h->is_top);
        h->h_top_list.clear(); h->h_top_by_level.assign(S.nlevels, {});
        for (int lv = 0; lv < S.nlevels; ++lv)
            for (int32_t q = S.level_ptr[lv]; q < S.level_ptr[lv + 1]; ++q) {
                const int32_t sn = S.level_list[q];
                if (h->is_top[sn]) { h->h_top_list.push_back(sn); h->h_top_by_level[lv].push_back(sn); }
            }
        std::vector<int8_t> active(S.nsuper), keep(S.N), topkeep(S.N);
        for (int32_t sn = 0; sn < S.nsuper; ++sn) {
            const bool top = h->is_top[sn];
            active[sn] = top ? (rank == 0) : (h->owner[sn] == rank);
            const int8_t kc = top ? (rank == 0) : (h->owner[sn] == rank);
            for (int32_t j = S.sn_first[sn]; j < S.sn_first[sn + 1]; ++j) {
                keep[j] = kc;
                topkeep[j] = top ? (rank == 0) : 1;
            }
        }
        cudaStream_t s = h->stream;
        CUDA_OK(h->d_active.upload(active, s)); CUDA_OK(h->d_keepcol.upload(keep, s));
        CUDA_OK(h->d_topcolkeep.upload(topkeep, s)); CUDA_OK(h->d_top_list.upload(h->h_top_list, s));
        // update blocks of supernodes other ranks own are never written here: keep them zero
        CUDA_OK(cudaMemsetAsync(h->d_U.p, 0, h->d_U.n * sizeof(double), s));
        int rc = build_plans(h);
        if (rc) return rc;
        // the captured launch sequences (if any) belong to the single-GPU plans
        for (GraphExec* g : {&h->g_factor[0], &h->g_factor[1], &h->g_solve}) {
            if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; }
            g->failed = false;
        }
        const char* env = getenv("CB200_DIST_GRAPH");
        if (!(env && env[0] == '1')) h->st.use_cuda_graph = 0;   // NCCL inside stream capture: opt-in
        CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    } catch (const std::exception& e) { set_error(e.what()); return -1; }
}

int32_t cb200_set_detail(cb200_handle* h, int32_t level) { h->detail = level < 0 ? 0 : (level > 2 ? 2 : level); return 0; }

static const char* const kFineNames[Timers::NFINE] = {
    "factor_small_nf16", "factor_small_nf32", "factor_small_nf64", "factor_small_nf96", "factor_small_nf128",
    "factor_small_nf152", "factor_panel_c0", "factor_panel_c1", "factor_panel_c2", "factor_panel_c3",
    "zero_update_blocks", "assemble_large", "piv_diag", "piv_rows", "schur_gemm", "finish_large", "factor_prologue",
    "fwd_leaf", "fwd_sub", "fwd_warp", "fwd_cta", "fwd_big_assemble", "fwd_big_tri", "fwd_big_gemv",
    "bwd_leaf", "bwd_sub", "bwd_warp", "bwd_cta", "bwd_big_gemvT", "bwd_big_tri", "permute_vectors", "nccl",
    "panel_update"};

int32_t cb200_get_fine_timers(cb200_handle* h, double* out_ms, int32_t len) {
    for (int i = 0; i < len && i < Timers::NFINE; ++i) out_ms[i] = h->tm.ms[Timers::NCOARSE + i];
    return Timers::NFINE;
}
const char* cb200_fine_timer_name(int32_t i) { return (i >= 0 && i < Timers::NFINE) ? kFineNames[i] : ""; }

int32_t cb200_get_stats(const cb200_handle* h, double* out, int32_t len) {
    const Symbolic& S = h->S;
    double schur = 0, panel_large = 0, big_bytes = 0, nlarge = 0;
    double my_flops = 0;
    for (int32_t sn = 0; sn < S.nsuper; ++sn) {
        const double ns = S.ns(sn), nr = S.nr(sn), nf = ns + nr;
        // multi-GPU: only the fronts THIS rank factors (its own subtrees + the replicated top)
        if (h->dist && !h->is_top[sn] && h->owner[sn] != h->rank) continue;
        for (double k = 0; k < ns; ++k) my_flops += (nf - k) * (nf - k);
        if (h->to_large((int)nf, (int)ns)) {
            nlarge += 1;
            schur += nr * (nr + 1.0) * ns;                 // flops of F22 -= L21 D L21' (lower part)
            for (double k = 0; k < ns; ++k) panel_large += (nf - k) * (nf - k);
        }
        if (nf * ns >= 65536 && ns > 32) big_bytes += 8.0 * nf * ns;
    }
    double v[14] = {S.flops, schur, panel_large - schur, (double)S.nnzL, (double)S.nlevels, (double)S.nsuper,
                    nlarge, big_bytes, (double)S.upd_total * 8.0, (double)S.panel_off.back() * 8.0,
                    (double)S.ordering_used, my_flops, h->use_tma ? 1.0 : 0.0, (double)h->tma_kmajor};
    for (int i = 0; i < len && i < 14; ++i) out[i] = v[i];
    return 0;
}

void* cb200_get_stream(cb200_handle* h) { return (void*)h->stream; }
int32_t cb200_set_resident(cb200_handle* h, int32_t resident) { h->resident = resident != 0; return 0; }
int32_t cb200_reset_timers(cb200_handle* h) {
    for (double& x : h->tm.ms) x = 0;
    h->tm.nfactor = h->tm.nsolve = h->tm.nlaunch = 0;
    return 0;
}

}  // extern "C"
