// api.cu -- C ABI of libgptpinn_b200.so (include/gptpinn_b200.h): argument checking, grouping of
// the parameter grid into warp work items, packing, launch.  No CPU fallback: a device that is
// not sm_100 is rejected with GPTPINN_E_UNSUPPORTED.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "../../include/gptpinn_b200.h"
#include "pack.cuh"
#include "precomp.cuh"
#include "pinn_train.cuh"
#include "sweep_kernel.cuh"
#include "sweep_rows.cuh"

namespace gp {
#define GP_DECL(P, N) cudaError_t sweep_launch_##P##_##N(const SweepParams &, int, int, int, size_t, cudaStream_t);
#define GP_DECL_ALL(P) GP_DECL(P, 1) GP_DECL(P, 2) GP_DECL(P, 3) GP_DECL(P, 4) GP_DECL(P, 5) GP_DECL(P, 6) GP_DECL(P, 7) GP_DECL(P, 8) GP_DECL(P, 9)
GP_DECL_ALL(0) GP_DECL_ALL(1) GP_DECL_ALL(2)
GP_DECL(0, 10) GP_DECL(0, 11) GP_DECL(0, 12) GP_DECL(0, 13) GP_DECL(0, 14) GP_DECL(0, 15)
#define GP_ROW9(P) sweep_launch_##P##_1, sweep_launch_##P##_2, sweep_launch_##P##_3, sweep_launch_##P##_4, sweep_launch_##P##_5, sweep_launch_##P##_6, sweep_launch_##P##_7, sweep_launch_##P##_8, sweep_launch_##P##_9
static const sweep_launch_fn kLaunch[3][15] = {
    {GP_ROW9(0), sweep_launch_0_10, sweep_launch_0_11, sweep_launch_0_12, sweep_launch_0_13, sweep_launch_0_14, sweep_launch_0_15},
    {GP_ROW9(1), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr},
    {GP_ROW9(2), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}};
#define GP_RDECL(P, N) cudaError_t sweep_rows_launch_##P##_##N(const RowsParams &, int, size_t, cudaStream_t); int sweep_rows_clusters_##P##_##N(int, size_t);
#define GP_RDECL_ALL(P) GP_RDECL(P, 1) GP_RDECL(P, 2) GP_RDECL(P, 3) GP_RDECL(P, 4) GP_RDECL(P, 5) GP_RDECL(P, 6) GP_RDECL(P, 7) GP_RDECL(P, 8) GP_RDECL(P, 9)
GP_RDECL_ALL(0) GP_RDECL_ALL(1) GP_RDECL_ALL(2)
GP_RDECL(0, 10) GP_RDECL(0, 11) GP_RDECL(0, 12) GP_RDECL(0, 13) GP_RDECL(0, 14) GP_RDECL(0, 15)
#define GP_RROW9(P) sweep_rows_launch_##P##_1, sweep_rows_launch_##P##_2, sweep_rows_launch_##P##_3, sweep_rows_launch_##P##_4, sweep_rows_launch_##P##_5, sweep_rows_launch_##P##_6, sweep_rows_launch_##P##_7, sweep_rows_launch_##P##_8, sweep_rows_launch_##P##_9
#define GP_CROW9(P) sweep_rows_clusters_##P##_1, sweep_rows_clusters_##P##_2, sweep_rows_clusters_##P##_3, sweep_rows_clusters_##P##_4, sweep_rows_clusters_##P##_5, sweep_rows_clusters_##P##_6, sweep_rows_clusters_##P##_7, sweep_rows_clusters_##P##_8, sweep_rows_clusters_##P##_9
static const rows_clusters_fn kRowsClusters[3][15] = {
    {GP_CROW9(0), sweep_rows_clusters_0_10, sweep_rows_clusters_0_11, sweep_rows_clusters_0_12, sweep_rows_clusters_0_13, sweep_rows_clusters_0_14, sweep_rows_clusters_0_15},
    {GP_CROW9(1), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr},
    {GP_CROW9(2), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}};
static const rows_launch_fn kRowsLaunch[3][15] = {
    {GP_RROW9(0), sweep_rows_launch_0_10, sweep_rows_launch_0_11, sweep_rows_launch_0_12, sweep_rows_launch_0_13, sweep_rows_launch_0_14, sweep_rows_launch_0_15},
    {GP_RROW9(1), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr},
    {GP_RROW9(2), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}};
static const int kMaxN[3] = {15, 9, 9};
}  // namespace gp

using namespace gp;

static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_TRY(x)                                                                                   \
    do {                                                                                              \
        cudaError_t e__ = (x);                                                                        \
        if (e__ != cudaSuccess) return fail(GPTPINN_E_CUDA, "%s failed: %s", #x, cudaGetErrorString(e__)); \
    } while (0)

struct gptpinn_handle {
    int device = -1;
    int n_sm = 0;
    float *packed = nullptr;   size_t packed_cap = 0;     // floats
    void *items = nullptr;     size_t items_cap = 0;      // bytes
    void *sibs = nullptr;      size_t sibs_cap = 0;       // bytes
    unsigned int *counter = nullptr;                      // work-item ticket
    void *train_ws = nullptr;  size_t train_cap = 0;      // packed params | adam m | adam v | partials | status | stop
    void *rows_ws = nullptr;   size_t rows_cap = 0;       // row-partitioned sweep: parameter table | coefficients | CTA partials
    void *tglob = nullptr;     size_t tglob_cap = 0;      // precomputed T arrays of the (alpha, beta) groups | their (tp0, tp1) table
};

static int ensure(void **ptr, size_t *cap, size_t need)
{
    if (need <= *cap) return GPTPINN_OK;
    if (*ptr) { cudaError_t e = cudaFree(*ptr); *ptr = nullptr; *cap = 0; if (e != cudaSuccess) return fail(GPTPINN_E_CUDA, "cudaFree: %s", cudaGetErrorString(e)); }
    size_t want = need + need / 4 + 4096;
    cudaError_t e = cudaMalloc(ptr, want);
    if (e != cudaSuccess) return fail(GPTPINN_E_NOMEM, "cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
    *cap = want;
    return GPTPINN_OK;
}

extern "C" int gptpinn_version(void) { return 200; }
extern "C" const char *gptpinn_last_error(void) { return g_err; }

extern "C" int gptpinn_create(gptpinn_handle **out)
{
    if (!out) return fail(GPTPINN_E_INVALID, "gptpinn_create: out is NULL");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        return fail(GPTPINN_E_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only (no fallback)",
                    dev, prop.major, prop.minor);
    gptpinn_handle *h = new (std::nothrow) gptpinn_handle();
    if (!h) return fail(GPTPINN_E_NOMEM, "out of host memory");
    h->device = dev;
    h->n_sm = prop.multiProcessorCount;
    *out = h;
    return GPTPINN_OK;
}

extern "C" int gptpinn_destroy(gptpinn_handle *h)
{
    if (!h) return GPTPINN_OK;
    if (h->packed) cudaFree(h->packed);
    if (h->items) cudaFree(h->items);
    if (h->sibs) cudaFree(h->sibs);
    if (h->counter) cudaFree(h->counter);
    if (h->train_ws) cudaFree(h->train_ws);
    if (h->rows_ws) cudaFree(h->rows_ws);
    if (h->tglob) cudaFree(h->tglob);
    delete h;
    return GPTPINN_OK;
}

// ------------------------------------------------------------------------------------------
// grid -> work items
// ------------------------------------------------------------------------------------------
struct ParamRow { float tp0, tp1, sp0, sp1; int idx; };

static inline uint64_t key_of(const ParamRow &r)
{
    uint32_t a, b;
    memcpy(&a, &r.tp0, 4);
    memcpy(&b, &r.tp1, 4);
    return ((uint64_t)a << 32) | b;
}

struct Plan {
    int M = 1, SL = 1, W = 1, ctas = 1, rounds = 1, groups = 0, priv = 0, K = 1;
    bool rows = false;                                   // row-partitioned kernel (sweep_rows.cuh) instead of work items
    double cost = 0.0;                                   // modelled SMSP cycles per epoch of the chosen work-item plan
    std::vector<ItemDesc> items;
    std::vector<SibDesc> sibs;
};

// ---- cost model (warp-instruction counts; one SMSP issues one warp-instruction per cycle) ----
static double per_row_instr(int pde, int n, int M)
{
    const double fma = (pde == PDE_B ? 6.0 : 4.0) * n;
    return M * (fma + 8.0) + (pde == PDE_B ? 3.0 : 2.0) * ((n + 3) / 4) + 8.0;
}
// set by run_sweep before planning: the private-ring kernel will find precomputed T arrays (no T build of any kind there)
static thread_local bool g_plan_tglob = false;

static double item_instr(int pde, int n, int M, int SL, long long rows_total, long long tiles_pass, int rt, int K = 1, bool shared_ring = false)
{
    const int RL = 32 / SL;
    const long long tiles = (tiles_pass + K - 1) / K;                  // a team of K warps splits the tiles of a pass
    const double rows_lane = (double)tiles * ((rt + RL - 1) / RL);     // every tile costs whole lane-rows
    // end-of-pass team sum: per parameter slot 2 named barriers + (n/4+1) stores and K*(n/4+1) loads + adds
    const double team = K > 1 ? M * (140.0 + (n / 4 + 1) * (2.0 + 2.0 * K) + (n + 1.0) * (K - 1)) : 0.0;
    // T: items with <= GP_REGT_MAX_SL sibling lanes form their T rows in registers (per row: extra loads + 2n FMA),
    // wider items build a per-warp T block in shared memory once per tile; plus ring bookkeeping per tile
    // per-tile ring cost in instruction-equivalents (mbarrier wait + first-row load latency, measured on B200)
    const double ring = shared_ring ? 300.0 : 200.0;
    const double build = (g_plan_tglob && !shared_ring) ? (double)tiles * ring
                       : SL <= GP_REGT_MAX_SL ? (double)tiles * ring + rows_lane * (2.0 * n + 6.0)
                                              : (double)tiles * (36.0 * (n / 4 + 1) * rt / 32.0 + ring);
    (void)rows_total;
    // measured issue efficiency relative to M = 5 (KG n = 15 single-round runs: 0.595 / 0.526 / 0.453 / 0.304 of FP32 peak)
    const double rel = M >= 5 ? 1.0 : (M == 4 ? 0.895 : (M >= 2 ? 0.835 : 0.624));
    return (rows_lane * per_row_instr(pde, n, M) + build + team) / rel;
}
// measured issue efficiency of one SMSP vs resident warps (B200, KG n=15: 1 warp 0.49/0.60, 2 warps 1.0 relative)
static double smsp_eff(double warps_per_smsp)
{
    if (warps_per_smsp <= 1.0) return 0.58;
    if (warps_per_smsp <= 2.0) return 0.58 + 0.15 * (warps_per_smsp - 1.0);
    if (warps_per_smsp <= 4.0) return 0.73 + 0.06 * (warps_per_smsp - 2.0);
    return 0.85;
}
static const double kL2BytesPerCycle = 4500.0;       // chip-wide L2 -> SM budget the planner allows itself

// makespan of items handed out in list order (= ticket order) to `workers` identical workers, each item going to
// the worker that frees up first.  Item times come in a few size classes, so worker loads are kept as a
// (load -> number of workers) map and whole runs of equal items are assigned at once.
static double makespan(const std::vector<double> &t, int workers)
{
    if (t.empty()) return 0.0;
    std::map<double, long long> loads;
    loads[0.0] = workers;
    double mx = 0.0;
    size_t i = 0;
    while (i < t.size()) {
        size_t j = i;
        while (j < t.size() && t[j] == t[i]) ++j;          // run of equal item times
        long long n = (long long)(j - i);
        const double s = t[i];
        while (n > 0) {
            auto it = loads.begin();                        // least-loaded workers
            const double L = it->first;
            const long long k = it->second;
            const long long m = n < k ? n : k;
            if (m == k) loads.erase(it); else it->second = k - m;
            loads[L + s] += m;
            if (L + s > mx) mx = L + s;
            n -= m;
        }
        i = j;
    }
    return mx;
}

struct Chunk { int g, s0, cnt, SL; };     // siblings [s0, s0+cnt) of group g on SL sibling-lanes

// binary decomposition of every group into power-of-two lane counts, largest first
static void decompose(const std::vector<int> &gstart, int M, int sl_cap, std::vector<Chunk> &out)
{
    out.clear();
    const int ngroups = (int)gstart.size() - 1;
    for (int g = 0; g < ngroups; ++g) {
        int S = gstart[g + 1] - gstart[g], s0 = 0;
        while (S > 0) {
            int lanes = (S + M - 1) / M;                 // sibling lanes still needed
            int SL = sl_cap;
            while (SL > lanes) SL >>= 1;                 // largest power of two <= lanes (and <= cap)
            if (SL < lanes && SL * 2 <= sl_cap && (long long)SL * 2 * M - S < M) SL *= 2;   // a last partial lane
            const int cnt = std::min(S, SL * M);
            out.push_back({g, s0, cnt, SL});
            s0 += cnt; S -= cnt;
        }
    }
    std::stable_sort(out.begin(), out.end(), [](const Chunk &x, const Chunk &y) { return x.SL > y.SL; });
}

// Tail refinement.  Tickets are handed out in list order (largest items first).  The items that
// do not fill a whole "round" of `workers` warps would leave most warps idle for a full item time,
// so that tail is cut into 2^k-times smaller items (k chosen by the cost model): the last warps
// to finish then differ by one *small* item.
static void refine_tail(std::vector<Chunk> &ch, int M, int workers, int pde, int n, long long tiles, int rt, int K = 1)
{
    if (ch.empty()) return;
    auto cost_of = [&](const std::vector<Chunk> &v) {
        std::vector<double> t(v.size());
        for (size_t i = 0; i < v.size(); ++i) t[i] = item_instr(pde, n, M, v[i].SL, 0, tiles, rt, K);
        return makespan(t, workers);
    };
    const size_t full = (ch.size() / (size_t)workers) * (size_t)workers;     // items in whole rounds keep their size
    if (full == ch.size()) return;
    std::vector<Chunk> head(ch.begin(), ch.begin() + full), tail(ch.begin() + full, ch.end());
    std::vector<Chunk> best = ch;
    double bestc = cost_of(ch);
    for (int k = 1; k <= 5; ++k) {
        std::vector<Chunk> nt;
        bool changed = false;
        for (const Chunk &c : tail) {
            if (c.SL <= 1) { nt.push_back(c); continue; }
            const int half = c.SL / 2, cap = half * M;
            const int c1 = std::min(c.cnt, cap);
            nt.push_back({c.g, c.s0, c1, half});
            if (c.cnt > c1) nt.push_back({c.g, c.s0 + c1, c.cnt - c1, half});
            changed = true;
        }
        if (!changed) break;
        std::stable_sort(nt.begin(), nt.end(), [](const Chunk &x, const Chunk &y) { return x.SL > y.SL; });
        tail.swap(nt);
        std::vector<Chunk> v(head);
        v.insert(v.end(), tail.begin(), tail.end());
        const double c = cost_of(v);
        if (c < bestc) { bestc = c; best.swap(v); }
    }
    ch.swap(best);
}

static int make_plan(std::vector<ParamRow> &rows, int pde, int n, int n_sm, const long long *seg_rows, int nseg,
                     const gptpinn_sweep_args *a, Plan &plan)
{
    std::stable_sort(rows.begin(), rows.end(), [](const ParamRow &x, const ParamRow &y) { return key_of(x) < key_of(y); });
    std::vector<int> gstart;
    for (size_t i = 0; i < rows.size(); ++i)
        if (i == 0 || key_of(rows[i]) != key_of(rows[i - 1])) gstart.push_back((int)i);
    gstart.push_back((int)rows.size());
    plan.groups = (int)gstart.size() - 1;

    long long tiles_p = 0, tiles_s = 0;
    for (int s = 0; s < nseg; ++s) { tiles_p += (seg_rows[s] + RT_PRIV - 1) / RT_PRIV; tiles_s += (seg_rows[s] + RT_SHARED - 1) / RT_SHARED; }
    const double slot_p = (pde == PDE_KG ? SlotFloats<PDE_KG, RT_PRIV>::value : pde == PDE_B ? SlotFloats<PDE_B, RT_PRIV>::value : SlotFloats<PDE_AC, RT_PRIV>::value) * 4.0;
    const double slot_s = (pde == PDE_KG ? SlotFloats<PDE_KG, RT_SHARED>::value : pde == PDE_B ? SlotFloats<PDE_B, RT_SHARED>::value : SlotFloats<PDE_AC, RT_SHARED>::value) * 4.0;

    static const int Ms[4] = {5, 4, 2, 1};
    int mode = a->cfg_mode;                           // 0 auto, 1 shared ring, 2 private rings
    if (a->bound_dev) {                               // early exit leaves the stream at a pass boundary: private rings only
        if (mode == 1) return fail(GPTPINN_E_INVALID, "bound_dev (early exit) needs the private-ring kernel (cfg_mode 0 or 2)");
        mode = 2;
    }
    if (a->cfg_M != 0 && !(a->cfg_M == 1 || a->cfg_M == 2 || a->cfg_M == 4 || a->cfg_M == 5))
        return fail(GPTPINN_E_INVALID, "cfg_M must be 0, 1, 2, 4 or 5");
    if (a->cfg_SL != 0 && (a->cfg_SL < 1 || a->cfg_SL > 32 || (a->cfg_SL & (a->cfg_SL - 1))))
        return fail(GPTPINN_E_INVALID, "cfg_SL must be 0 or a power of two <= 32");
    if (a->cfg_mode < 0 || a->cfg_mode > 4) return fail(GPTPINN_E_INVALID, "cfg_mode must be 0 .. 4");

    if (a->cfg_K != 0 && !(a->cfg_K == 1 || a->cfg_K == 2 || a->cfg_K == 4 || a->cfg_K == 8))
        return fail(GPTPINN_E_INVALID, "cfg_K must be 0, 1, 2, 4 or 8");
    double best = 1e300;
    int bM = 0, bSL = 0, bW = 0, bPriv = 0, bK = 1;
    std::vector<Chunk> bestChunks, ch;
    int maxS = 1;
    for (int g = 0; g < plan.groups; ++g) maxS = std::max(maxS, gstart[g + 1] - gstart[g]);
    int Mcap = 5;                                       // parameters per thread beyond the largest sibling group only idle
    for (int mi = 3; mi >= 0; --mi) if (Ms[mi] >= maxS) { Mcap = Ms[mi]; break; }
    for (int mi = 0; mi < 4; ++mi) {
        const int M = Ms[mi];
        if (a->cfg_M > 0 && M != a->cfg_M) continue;
        if (a->cfg_M == 0 && M > Mcap) continue;
        const int maxw = maxw_for_m(M);
        if (a->cfg_M == 0 && best < 1e300) {
            // lower bound for this M: every sibling lane at the per-lane cost of the widest item, perfectly spread over all
            // warps an SM can hold -- if even that loses to the best plan so far, skip the search
            double lanes = 0.0;
            for (int g = 0; g < plan.groups; ++g) lanes += (double)((gstart[g + 1] - gstart[g] + M - 1) / M);
            const double per_lane = std::min(item_instr(pde, n, M, 32, 0, tiles_p, RT_PRIV) / 32.0, item_instr(pde, n, M, 32, 0, tiles_s, RT_SHARED, 1, true) / 32.0);
            const double wps = maxw / 4.0;
            const double lb = lanes * per_lane / ((double)n_sm * maxw) * std::ceil(wps) / smsp_eff(wps);
            if (lb >= best) continue;
        }
        // ---- private rings: dynamic tickets, mixed item sizes ----
        if (mode != 1) {
            const int wsmem = (int)((227.0 * 1024 - 96) / (NST_PRIV * slot_p + NST_PRIV * 8 + TEAM_BYTES_PER_WARP));
            const int maxw1 = std::max(1, std::min(maxw, wsmem));
          for (int K = 1; K <= 8; K <<= 1) {
            // teams of K warps share an item (finer time quanta for grids that give each warp only a few items)
            if (a->cfg_K > 0 && K != a->cfg_K) continue;
            if (K > maxw1 || K > tiles_p) continue;
            const int maxwp = (maxw1 / K) * K;
            for (int cap = 32; cap >= 1; cap >>= 1) {
                std::vector<Chunk> base;
                if (a->cfg_SL > 0) {
                    if (cap != 32) break;
                    for (int g = 0; g < plan.groups; ++g)
                        for (int s0 = 0, S = gstart[g + 1] - gstart[g]; s0 < S; s0 += a->cfg_SL * M)
                            base.push_back({g, s0, std::min(S - s0, a->cfg_SL * M), a->cfg_SL});
                } else {
                    decompose(gstart, M, cap, base);
                    if (cap < 32 && !base.empty() && base.front().SL < cap) continue;  // the cap does not bind: same as the previous one
                    if ((long long)base.size() * K > 12LL * n_sm * maxwp) break;       // finer than 12 items per resident team never pays
                }
                // with at least one item per resident warp the full warp count is always best
                const int wlo = (a->cfg_W == 0 && (long long)base.size() * K >= (long long)n_sm * maxwp) ? maxwp : std::max(K, maxwp - 3);
                for (int W = maxwp; W >= wlo; W -= K) {
                    if (a->cfg_W > 0 && W != std::max(K, std::min(a->cfg_W, maxwp) / K * K)) continue;
                    const int TPC = W / K;                                  // teams per CTA
                    ch = base;
                    if (a->cfg_SL == 0) refine_tail(ch, M, n_sm * TPC, pde, n, tiles_p, RT_PRIV, K);
                    std::vector<double> t(ch.size());
                    for (size_t i = 0; i < ch.size(); ++i) t[i] = item_instr(pde, n, M, ch[i].SL, 0, tiles_p, RT_PRIV, K);
                    const int Tuse = (int)std::max<long long>(1, std::min<long long>(TPC, ((long long)ch.size() + n_sm - 1) / n_sm));
                    const double wps = Tuse * K / 4.0;
                    const double issue = makespan(t, n_sm * Tuse) * std::max(1.0, std::ceil(wps)) / smsp_eff(std::max(1.0, wps));
                    const double l2 = (double)ch.size() * tiles_p * slot_p / kL2BytesPerCycle;
                    const double cst = std::max(issue, l2) * (1.0 + 0.005 * (K > 1));      // ties go to independent warps
                    if (cst < best) { best = cst; bM = M; bSL = ch.empty() ? 1 : ch.front().SL; bW = Tuse * K; bPriv = 1; bK = K; bestChunks = ch; }
                }
            }
          }
        }
        // ---- shared ring: static rounds, uniform item size ----
        if (mode != 2) {
            for (int SL = 1; SL <= 32; SL <<= 1) {
                if (a->cfg_SL > 0 && SL != a->cfg_SL) continue;
                long long nit = 0;
                for (int g = 0; g < plan.groups; ++g) { const int S = gstart[g + 1] - gstart[g]; nit += (S + M * SL - 1) / (M * SL); }
                for (int W = 1; W <= maxw; ++W) {
                    if (a->cfg_W > 0 && W != std::min(a->cfg_W, maxw)) continue;
                    const long long ctas = std::min<long long>(n_sm, (nit + W - 1) / W);
                    const long long full_rounds = nit / (ctas * W), rest = nit - full_rounds * ctas * W;
                    const double rounds = (double)(full_rounds + (rest > 0 ? 1 : 0));
                    const double t_item = item_instr(pde, n, M, SL, 0, tiles_s, RT_SHARED, 1, true);
                    double issue = full_rounds * std::ceil(W / 4.0) * t_item / smsp_eff(W / 4.0);
                    if (rest > 0) {                                         // the last round runs with fewer warps per SM
                        const double wl = (double)((rest + ctas - 1) / ctas);
                        issue += std::ceil(wl / 4.0) * t_item / smsp_eff(std::max(1.0, wl / 4.0));
                    }
                    const double l2 = rounds * ctas * tiles_s * slot_s / kL2BytesPerCycle;
                    const double cst = std::max(issue, l2) * 1.02;          // ties go to the private rings
                    if (cst < best) { best = cst; bM = M; bSL = SL; bW = W; bPriv = 0; bK = 1; }
                }
            }
        }
    }
    if (bM == 0) return fail(GPTPINN_E_INVALID, "no feasible sweep configuration");
    plan.cost = best;
    plan.M = bM; plan.SL = bSL; plan.W = std::max(1, bW); plan.priv = bPriv; plan.K = bK;
    if (!bPriv) {
        bestChunks.clear();
        for (int g = 0; g < plan.groups; ++g)
            for (int s0 = 0, S = gstart[g + 1] - gstart[g]; s0 < S; s0 += bSL * bM)
                bestChunks.push_back({g, s0, std::min(S - s0, bSL * bM), bSL});
    }
    plan.items.clear(); plan.sibs.clear();
    plan.sibs.reserve(rows.size());
    plan.items.reserve(bestChunks.size());
    for (const Chunk &c : bestChunks) {
        const int base = gstart[c.g] + c.s0;
        ItemDesc it;
        memset(&it, 0, sizeof(it));
        it.t_slot = -1;
        it.tp0 = rows[base].tp0; it.tp1 = rows[base].tp1;
        it.sib_start = (int)plan.sibs.size();
        it.sib_count = c.cnt;
        it.SL = c.SL;
        for (int s = 0; s < c.cnt; ++s) {
            SibDesc sd;
            sd.sp0 = rows[base + s].sp0; sd.sp1 = rows[base + s].sp1; sd.out_idx = rows[base + s].idx; sd.pad = 0;
            plan.sibs.push_back(sd);
        }
        plan.items.push_back(it);
    }
    const long long nit = (long long)plan.items.size();
    if (bPriv) {
        const int tpc = plan.W / plan.K;
        plan.ctas = (int)std::min<long long>(n_sm, (nit + tpc - 1) / tpc);
        plan.rounds = (int)((nit + (long long)plan.ctas * tpc - 1) / ((long long)plan.ctas * tpc));
    } else {
        plan.ctas = (int)std::min<long long>(n_sm, (nit + plan.W - 1) / plan.W);
        plan.rounds = (int)((nit + (long long)plan.ctas * plan.W - 1) / ((long long)plan.ctas * plan.W));
    }
    return GPTPINN_OK;
}

// ------------------------------------------------------------------------------------------
// common driver
// ------------------------------------------------------------------------------------------
struct Segments { long long rows[MAX_SEG]; int nseg; };

// ------------------------------------------------------------------------------------------
// row-partitioned kernel (sweep_rows.cuh): cost model and launch
// ------------------------------------------------------------------------------------------
struct RowsGeom { int npsp, rs, chunk_rows; size_t smem; bool ok; int S, C; };

// G = CTAs that share the rows (the whole grid, or the S CTAs of a cluster); npar = parameters every such CTA works on
// (all of them, or the cluster's share); cluster flavour (S > 0): every share must be resident
static RowsGeom rows_geometry(int pde, int n, int npar, const Segments &seg, int G, int chunk_override = 0, int S = 0, int C = 0)
{
    RowsGeom g = {1, ROWS_THREADS, 0, 0, false, S, C};
    if (npar < 1 || npar > ROWS_MAX_PARAMS) return g;
    const int nps = (npar + ROWS_M - 1) / ROWS_M;
    g.npsp = nps; g.rs = ROWS_THREADS / nps;
    const int narr = pde == PDE_B ? 4 : 3, nvec = pde == PDE_KG ? 2 : 1, nqx = (n + 3) / 4;
    const size_t row_floats = (size_t)narr * nqx * 4 + nvec;
    size_t fl = 0;
    for (int s = 1; s < seg.nseg; ++s) {
        long long share = (seg.rows[s] + G - 1) / G + 1;
        share = std::max<long long>(4, (share + 3) & ~3LL);
        fl += (size_t)share * row_floats + 4;
    }
    // slice-reduction buffer [rs][npsp * M][ROWS_KC] (rs > 1) and, aliased with it (grid flavour), the staged coefficient
    // table [n][npsp * M]; cluster flavour: + this CTA's partial sums and its copy of the coefficient table
    if (S > 0) fl += (g.rs > 1 ? (size_t)ROWS_THREADS * ROWS_M * ROWS_KC : 0) + (((size_t)nps * ROWS_M * (n + 1) + 3) & ~(size_t)3) + (size_t)nps * ROWS_M * n;
    else fl += std::max(g.rs > 1 ? (size_t)ROWS_THREADS * ROWS_M * ROWS_KC : 0, (size_t)nps * ROWS_M * n);
    const size_t budget = (size_t)(220 * 1024) / sizeof(float);
    if (fl + 64 * row_floats > budget) return g;
    long long share0 = (seg.rows[0] + G - 1) / G + 1;
    share0 = std::max<long long>(4, (share0 + 3) & ~3LL);
    const long long cap_total = (long long)((budget - fl) / row_floats) & ~3LL;
    int nbuf = 1;
    if (share0 <= cap_total) g.chunk_rows = (int)share0;                // the CTA's residual share stays resident
    else if (S > 0) return g;
    else { g.chunk_rows = (int)std::min<long long>((cap_total / 2) & ~3LL, 512); nbuf = 2; }   // streamed through a double buffer
    if (S == 0 && chunk_override > 0 && ((chunk_override + 3) & ~3) < g.chunk_rows) { g.chunk_rows = std::max(4, (chunk_override + 3) & ~3); nbuf = 2; }   // tests: force streaming
    g.smem = (fl + (size_t)nbuf * g.chunk_rows * row_floats + 8) * sizeof(float);
    g.ok = true;
    return g;
}

// modelled SMSP cycles per epoch, in the units of item_instr(): FMAs of the T-free formulation + row loads on every CTA's
// share of the rows for its parameters, plus the per-epoch exchange (calibrated on B200: the grid flavour pays two grid
// barriers and a global-memory round trip of the partials, about 9 us; the cluster flavour two cluster barriers and the
// distributed-shared-memory sums)
static double rows_cost(int pde, int n, int npar_total, const Segments &seg, int G, const RowsGeom &g)
{
    const double fma_res = (pde == PDE_B ? 8.0 : 6.0) * n + 14.0, fma_icbc = (pde == PDE_KG ? 4.0 : 2.0) * n + 8.0, fma_bc = (pde == PDE_AC ? 3.0 : 2.0) * n + 6.0;
    double rows_inst = 0.0;
    for (int s = 0; s < seg.nseg; ++s) {
        const double share = std::ceil(std::ceil((double)seg.rows[s] / G) / g.rs);            // rows per thread
        rows_inst += share * ROWS_M * (s == 0 ? fma_res : (s == 1 ? fma_icbc : fma_bc));
    }
    const double issue = rows_inst * 2.0 / 0.58;                        // 8 warps = 2 per SMSP, measured issue efficiency of this loop
    const int passes = (n + 1 + ROWS_KC - 1) / ROWS_KC;
    const double slices = g.rs > 1 ? passes * (g.rs * 8.0 + 400.0) : 0.0;
    // measured on B200 (tools/check_rows.py): the cluster flavour costs about 4.5 us per epoch besides the row loops (two
    // cluster barriers, the distributed-shared-memory sums, the table reload, loop start-up latencies) ...
    if (g.S > 0) return issue + slices + 7000.0;
    // ... the grid flavour two grid barriers (about 3.3 us each), the owner phase (G partials per pair, `tpp` lanes per
    // pair with four loads in flight: ceil(G / (4 tpp)) dependent L2 round trips) and the partial-sum traffic
    int tpp = 1;
    while (tpp < 32 && (long long)npar_total * (n + 1) * (tpp * 2) <= (long long)G * ROWS_THREADS) tpp *= 2;
    const double barriers = 2.0 * 6500.0;
    const double traffic = 2.0 * (double)npar_total * (n + 1) * 4.0 * G / kL2BytesPerCycle + std::ceil(G / (4.0 * tpp)) * 1500.0 + 2500.0;
    return issue + slices + barriers + traffic;
}

static int run_sweep_rows(gptpinn_handle *h, int pde, int n, const Segments &seg, PackJobs &jobs, bool has_fhat,
                          int NR, int NI, int NB, const std::vector<ParamRow> &rows, const gptpinn_sweep_args *a,
                          gptpinn_sweep_info *info, const RowsGeom &geo, int G, cudaStream_t st)
{
    const int RT = RT_PRIV;
    const int slot = pde == PDE_KG ? SlotFloats<PDE_KG, RT_PRIV>::value : (pde == PDE_B ? SlotFloats<PDE_B, RT_PRIV>::value : SlotFloats<PDE_AC, RT_PRIV>::value);
    RowsParams p;
    memset(&p, 0, sizeof(p));
    long long tiles = 0, max_rows = 0, tile0[MAX_SEG];
    for (int s = 0; s < seg.nseg; ++s) {
        tile0[s] = tiles;
        p.seg_rows[s] = (int)seg.rows[s];
        p.seg_tile0[s] = (int)tiles;
        tiles += (seg.rows[s] + RT - 1) / RT;
        max_rows = std::max(max_rows, seg.rows[s]);
    }
    for (int i = 0; i < jobs.nmat; ++i) jobs.mat[i].tile0 = tile0[jobs.mat[i].tile0];
    for (int i = 0; i < jobs.nvec; ++i) jobs.vec[i].tile0 = tile0[jobs.vec[i].tile0];
    jobs.slot = slot;
    jobs.rt = RT;
    size_t cap_f = h->packed_cap * sizeof(float);
    int rc = ensure((void **)&h->packed, &cap_f, (size_t)tiles * slot * sizeof(float));
    h->packed_cap = cap_f / sizeof(float);
    if (rc) return rc;
    const int Np = a->Np, NV = n + 1;
    const size_t pp_b = (size_t)Np * sizeof(float4), cg_b = (size_t)geo.npsp * ROWS_M * n * sizeof(float);      // [n][npsp * M]
    const size_t part_b = geo.S > 0 ? 16 : (size_t)G * Np * NV * sizeof(float);
    rc = ensure(&h->rows_ws, &h->rows_cap, pp_b + cg_b + part_b);
    if (rc) return rc;
    std::vector<float4> pp((size_t)Np);
    for (const ParamRow &r : rows) pp[(size_t)r.idx] = make_float4(r.tp0, r.tp1, r.sp0, r.sp1);
    CUDA_TRY(cudaMemsetAsync(h->packed, 0, (size_t)tiles * slot * sizeof(float), st));
    CUDA_TRY(launch_pack(jobs, h->packed, max_rows, st));
    CUDA_TRY(cudaMemcpyAsync(h->rows_ws, pp.data(), pp_b, cudaMemcpyHostToDevice, st));      // pageable source: consumed on return
    p.packed = h->packed;
    p.nseg = seg.nseg;
    p.slot = slot;
    p.chunk_rows = geo.chunk_rows;
    p.invR = (float)(1.0 / NR); p.invI = (float)(1.0 / NI); p.invB = (float)(1.0 / NB);
    p.fac1 = (float)(2.0 / NR); p.fac2 = (float)(2.0 / NB); p.fac3 = (float)(2.0 / NI);
    p.lr = (float)a->lr;
    p.epochs = a->epochs;
    p.rec_every = (a->trace_dev && a->rec_every > 0) ? a->rec_every : 0;
    p.nrec = p.rec_every > 0 ? a->epochs / p.rec_every + 1 : 0;
    p.has_fhat = has_fhat ? 1 : 0;
    p.Np = Np; p.npsp = geo.npsp; p.rs = geo.rs;
    p.S = geo.S; p.C = geo.C;
    p.tpp = 1;
    while (p.tpp < 32 && (long long)Np * NV * (p.tpp * 2) <= (long long)G * ROWS_THREADS) p.tpp *= 2;
    p.pp = (const float4 *)h->rows_ws;
    p.c0 = a->c0_dev; p.c0_per_param = a->c0_per_param;
    p.c_out = a->c_out_dev; p.f_out = a->f_out_dev; p.m_out = a->m_out_dev;
    p.trace = p.rec_every > 0 ? a->trace_dev : nullptr;
    p.key_out = a->key_out_dev; p.key_index = a->key_index_dev;
    p.cglob = (float *)((char *)h->rows_ws + pp_b);
    p.partial = (float *)((char *)h->rows_ws + pp_b + cg_b);
    rows_launch_fn fn = kRowsLaunch[pde][n - 1];
    if (!fn) return fail(GPTPINN_E_UNSUPPORTED, "no row-partitioned kernel compiled for pde=%d n=%d", pde, n);
    CUDA_TRY(fn(p, G, geo.smem, st));
    if (info) {
        memset(info, 0, sizeof(*info));
        info->M = ROWS_M; info->SL = geo.npsp; info->W = ROWS_THREADS / 32; info->items = Np; info->groups = Np;
        info->ctas = G; info->rounds = 1; info->priv = geo.S > 0 ? 3 : 2; info->K = geo.S > 0 ? geo.S : geo.rs; info->kernels_launched = 2;
        info->smem_bytes = (int)geo.smem; info->tiles_per_pass = (int)tiles;
    }
    return GPTPINN_OK;
}

static int run_sweep(gptpinn_handle *h, int pde, int n, const Segments &seg, PackJobs &jobs, bool has_fhat,
                     int NR, int NI, int NB, std::vector<ParamRow> &rows, const gptpinn_sweep_args *a,
                     gptpinn_sweep_info *info, cudaStream_t st)
{
    int dev = -1;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev != h->device) return fail(GPTPINN_E_INVALID, "handle belongs to device %d, current device is %d", h->device, dev);
    if (a->epochs < 0) return fail(GPTPINN_E_INVALID, "epochs < 0");
    if (a->Np > 0 && (!a->f_out_dev || !a->m_out_dev || !a->c0_dev)) return fail(GPTPINN_E_INVALID, "f_out_dev, m_out_dev and c0_dev are required");

    if (a->Np == 0) {
        if (info) memset(info, 0, sizeof(*info));
        return GPTPINN_OK;
    }
    // precomputed T arrays (one per (tp0, tp1) group) are used when they fit the memory budget and f_hat is zero; the
    // planner prices the work items accordingly
    size_t tglob_budget_mb = 8192;
    if (const char *ev = getenv("GPTPINN_TGLOB_MB")) tglob_budget_mb = (size_t)std::max(0, atoi(ev));
    struct FlagGuard { ~FlagGuard() { g_plan_tglob = false; } } flag_guard;       // the flag never outlives this call
    {
        std::vector<uint64_t> keys(rows.size());
        for (size_t i = 0; i < rows.size(); ++i) keys[i] = key_of(rows[i]);
        std::sort(keys.begin(), keys.end());
        const size_t ngroups = (size_t)(std::unique(keys.begin(), keys.end()) - keys.begin());
        const size_t per_group = (size_t)((seg.rows[0] + RT_PRIV - 1) / RT_PRIV) * (size_t)(n / 4 + 1) * 4 * RT_PRIV * sizeof(float);
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const size_t need = ngroups * per_group + ngroups * sizeof(float2) + 256;
        g_plan_tglob = !has_fhat && !getenv("GPTPINN_NO_TGLOB") && need <= tglob_budget_mb * 1024 * 1024 && need + need / 4 <= h->tglob_cap + free_b / 2;
    }
    // row-partitioned kernels (sweep_rows.cuh): few parameters, no early-exit bound.  cfg_mode 3 = grid flavour (cooperative
    // launch, any row count), cfg_mode 4 = cluster flavour (cfg_K = CTAs per cluster, 0 = best), cfg_mode 0 = cheapest of
    // the work-item plan and the two flavours by the cost model
    if (a->cfg_mode == 3 || a->cfg_mode == 4 || (a->cfg_mode == 0 && !a->bound_dev && !a->cfg_M && !a->cfg_SL && !a->cfg_W && !a->cfg_K)) {
        if (a->cfg_mode != 0 && a->bound_dev) return fail(GPTPINN_E_INVALID, "bound_dev (early exit) is not available in the row-partitioned kernels");
        int coop = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        RowsGeom best_geo = {1, ROWS_THREADS, 0, 0, false, 0, 0};
        double best_cost = 1e300;
        int best_G = 0;
        if (a->cfg_mode != 4 && coop && a->Np <= ROWS_MAX_PARAMS) {
            const RowsGeom geo = rows_geometry(pde, n, a->Np, seg, h->n_sm, a->cfg_mode == 3 ? a->cfg_W : 0);
            if (geo.ok) { best_geo = geo; best_cost = rows_cost(pde, n, a->Np, seg, h->n_sm, geo); best_G = h->n_sm; }
        }
        if (a->cfg_mode != 3 && kRowsClusters[pde][n - 1]) {
            for (int S = 16; S >= 1; S >>= 1) {
                if (a->cfg_mode == 4 && a->cfg_K > 0 && S != a->cfg_K) continue;
                int C = std::min(h->n_sm / S, (a->Np + ROWS_M - 1) / ROWS_M);          // at least one parameter set per cluster
                if (const char *ov = getenv("GPTPINN_ROWS_C")) C = std::max(1, std::min(C, atoi(ov)));   // experiments: fewer clusters
                if (C < 1) continue;
                if ((a->Np + C - 1) / C > ROWS_MAX_PARAMS) continue;
                RowsGeom geo = rows_geometry(pde, n, (a->Np + C - 1) / C, seg, S, 0, S, C);
                if (!geo.ok) continue;
                // co-resident clusters of this shape (driver query: cached per (device, pde, n, S, shared-memory size))
                static thread_local std::map<std::pair<long long, size_t>, int> fit_cache;
                const std::pair<long long, size_t> fkey(((long long)dev << 32) | (pde << 16) | (n << 8) | S, geo.smem);
                auto fc = fit_cache.find(fkey);
                if (fc == fit_cache.end()) fc = fit_cache.emplace(fkey, kRowsClusters[pde][n - 1](S, geo.smem)).first;
                const int fit = fc->second;
                if (fit < 1) continue;
                if (fit < C) {
                    C = fit;
                    if ((a->Np + C - 1) / C > ROWS_MAX_PARAMS) continue;
                    geo = rows_geometry(pde, n, (a->Np + C - 1) / C, seg, S, 0, S, C);
                    if (!geo.ok) continue;
                }
                const double cst = rows_cost(pde, n, a->Np, seg, S, geo);
                if (cst < best_cost) { best_cost = cst; best_geo = geo; best_G = S * C; }
            }
        }
        if (a->cfg_mode != 0) {
            if (!best_geo.ok) return fail(GPTPINN_E_UNSUPPORTED, "row-partitioned kernel: needs <= %d parameters per CTA, a cooperative / cluster launch and the row shares in shared memory", ROWS_MAX_PARAMS);
            g_plan_tglob = false;
            return run_sweep_rows(h, pde, n, seg, jobs, has_fhat, NR, NI, NB, rows, a, info, best_geo, best_G, st);
        }
        if (best_geo.ok) {
            Plan probe;
            std::vector<ParamRow> tmp(rows);
            int rcp = make_plan(tmp, pde, n, h->n_sm, seg.rows, seg.nseg, a, probe);
            if (rcp) return rcp;
            if (best_cost < probe.cost) {
                g_plan_tglob = false;
                return run_sweep_rows(h, pde, n, seg, jobs, has_fhat, NR, NI, NB, rows, a, info, best_geo, best_G, st);
            }
        }
    }
    Plan plan;
    int rc = make_plan(rows, pde, n, h->n_sm, seg.rows, seg.nseg, a, plan);
    const bool use_tglob = g_plan_tglob;
    g_plan_tglob = false;
    if (rc) return rc;

    const int RT = plan.priv ? RT_PRIV : RT_SHARED;
    const int slot = plan.priv ? (pde == PDE_KG ? SlotFloats<PDE_KG, RT_PRIV>::value : (pde == PDE_B ? SlotFloats<PDE_B, RT_PRIV>::value : SlotFloats<PDE_AC, RT_PRIV>::value))
                               : (pde == PDE_KG ? SlotFloats<PDE_KG, RT_SHARED>::value : (pde == PDE_B ? SlotFloats<PDE_B, RT_SHARED>::value : SlotFloats<PDE_AC, RT_SHARED>::value));
    SweepParams p;
    memset(&p, 0, sizeof(p));
    long long tiles = 0, max_rows = 0;
    long long tile0[MAX_SEG];
    for (int s = 0; s < seg.nseg; ++s) {
        tile0[s] = tiles;
        p.ntile[s] = (int)((seg.rows[s] + RT - 1) / RT);
        tiles += p.ntile[s];
        max_rows = std::max(max_rows, seg.rows[s]);
    }
    for (int i = 0; i < jobs.nmat; ++i) jobs.mat[i].tile0 = tile0[jobs.mat[i].tile0];
    for (int i = 0; i < jobs.nvec; ++i) jobs.vec[i].tile0 = tile0[jobs.vec[i].tile0];
    jobs.slot = slot;
    jobs.rt = RT;
    const unsigned long long total_tiles = (unsigned long long)plan.rounds * (unsigned long long)(a->epochs + 1) * (unsigned long long)tiles;
    if (!plan.priv && total_tiles >= 0xfffffff0ull) return fail(GPTPINN_E_UNSUPPORTED, "rounds*epochs*tiles = %llu overflows the tile counter", total_tiles);

    size_t cap_f = h->packed_cap * sizeof(float);
    rc = ensure((void **)&h->packed, &cap_f, (size_t)tiles * slot * sizeof(float));
    h->packed_cap = cap_f / sizeof(float);
    if (rc) return rc;
    rc = ensure(&h->items, &h->items_cap, plan.items.size() * sizeof(ItemDesc));
    if (rc) return rc;
    rc = ensure(&h->sibs, &h->sibs_cap, plan.sibs.size() * sizeof(SibDesc));
    if (rc) return rc;

    // ---- precomputed T: every group that has a wide work item (T block path of the private-ring kernel) gets its T array
    //      built once per launch in global memory; the kernel then fetches T tiles instead of rebuilding them per tile and
    //      epoch.  Skipped (in-kernel build) when the arrays exceed the memory budget or f_hat is non-zero.
    std::vector<float2> tgroups;
    const long long tt = (long long)(n / 4 + 1) * 4 * RT, t_stride = (long long)p.ntile[0] * tt;
    if (plan.priv && use_tglob) {
        std::map<uint64_t, int> slot_of;
        for (ItemDesc &it : plan.items) {
            uint32_t a_, b_;
            memcpy(&a_, &it.tp0, 4); memcpy(&b_, &it.tp1, 4);
            const uint64_t key = ((uint64_t)a_ << 32) | b_;
            auto f = slot_of.find(key);
            if (f == slot_of.end()) { f = slot_of.emplace(key, (int)tgroups.size()).first; tgroups.push_back(make_float2(it.tp0, it.tp1)); }
            it.t_slot = f->second;
        }
        const size_t need = (size_t)tgroups.size() * (size_t)t_stride * sizeof(float) + tgroups.size() * sizeof(float2) + 256;
        rc = ensure(&h->tglob, &h->tglob_cap, need);                            // (budget and free memory checked before planning)
        if (rc) return rc;
    }

    if (!h->counter) CUDA_TRY(cudaMalloc((void **)&h->counter, sizeof(unsigned int)));
    CUDA_TRY(cudaMemsetAsync(h->counter, 0, sizeof(unsigned int), st));
    CUDA_TRY(cudaMemsetAsync(h->packed, 0, (size_t)tiles * slot * sizeof(float), st));
    CUDA_TRY(launch_pack(jobs, h->packed, max_rows, st));
    if (!tgroups.empty()) {
        float2 *gtab = (float2 *)((char *)h->tglob + (((size_t)tgroups.size() * (size_t)t_stride * sizeof(float) + 255) & ~(size_t)255));
        CUDA_TRY(cudaMemcpyAsync(gtab, tgroups.data(), tgroups.size() * sizeof(float2), cudaMemcpyHostToDevice, st));
        CUDA_TRY(launch_build_t(h->packed, slot, RT, pde == PDE_B ? 4 : 3, pde, n, p.ntile[0], gtab, (int)tgroups.size(), (float *)h->tglob, t_stride, st));
        p.tglob = (const float *)h->tglob;
        p.t_stride = t_stride;
    }
    CUDA_TRY(cudaMemcpyAsync(h->items, plan.items.data(), plan.items.size() * sizeof(ItemDesc), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->sibs, plan.sibs.data(), plan.sibs.size() * sizeof(SibDesc), cudaMemcpyHostToDevice, st));
    // pageable sources: the copies above have consumed the host vectors when the calls return

    p.packed = h->packed;
    p.tiles_per_pass = (int)tiles;
    p.invR = (float)(1.0 / NR); p.invI = (float)(1.0 / NI); p.invB = (float)(1.0 / NB);
    p.fac1 = (float)(2.0 / NR); p.fac2 = (float)(2.0 / NB); p.fac3 = (float)(2.0 / NI);
    p.lr = (float)a->lr;
    p.epochs = a->epochs;
    p.rec_every = (a->trace_dev && a->rec_every > 0) ? a->rec_every : 0;
    p.nrec = p.rec_every > 0 ? a->epochs / p.rec_every + 1 : 0;
    p.has_fhat = has_fhat ? 1 : 0;
    p.items = (const ItemDesc *)h->items;
    p.n_items = (int)plan.items.size();
    p.sibs = (const SibDesc *)h->sibs;
    p.c0 = a->c0_dev;
    p.c0_per_param = a->c0_per_param;
    p.c_out = a->c_out_dev; p.f_out = a->f_out_dev; p.m_out = a->m_out_dev;
    p.trace = p.rec_every > 0 ? a->trace_dev : nullptr;
    p.W = plan.W;
    p.rounds = plan.rounds;
    p.counter = h->counter;
    p.K = plan.priv ? plan.K : 1;
    p.bound = plan.priv ? a->bound_dev : nullptr;
    p.key_out = a->key_out_dev; p.key_index = a->key_index_dev;

    const size_t tblk = (size_t)(n / 4 + 1) * 4 * RT * 4;                       // shared ring only: per-warp T block
    const size_t smem = plan.priv
        ? (size_t)plan.W * NST_PRIV * slot * 4 + (((size_t)plan.W * NST_PRIV * sizeof(uint64_t) + 15) & ~(size_t)15)
              + (size_t)plan.W * TEAM_BYTES_PER_WARP
        : (size_t)NSTAGE * slot * 4 + (size_t)plan.W * tblk + 2 * NSTAGE * sizeof(uint64_t) + 16;
    if (smem > 227 * 1024) return fail(GPTPINN_E_UNSUPPORTED, "configuration needs %zu bytes of shared memory", smem);
    sweep_launch_fn fn = kLaunch[pde][n - 1];
    if (!fn) return fail(GPTPINN_E_UNSUPPORTED, "no kernel compiled for pde=%d n=%d", pde, n);
    if (p.n_items > 0) CUDA_TRY(fn(p, plan.M, plan.priv, plan.ctas, smem, st));
    if (info) {
        info->M = plan.M; info->SL = plan.SL; info->W = plan.W; info->items = p.n_items; info->groups = plan.groups;
        info->ctas = plan.ctas; info->rounds = plan.rounds; info->priv = plan.priv; info->K = plan.priv ? plan.K : 1;
        info->kernels_launched = (p.n_items > 0 ? 2 : 1) + (tgroups.empty() ? 0 : 1);
        info->smem_bytes = (int)smem; info->tiles_per_pass = (int)tiles;
    }
    return GPTPINN_OK;
}

static int check_mat(const gptpinn_mat &m, int n, const char *name)
{
    if (!m.ptr) return fail(GPTPINN_E_INVALID, "%s is NULL", name);
    if (m.ld < n) return fail(GPTPINN_E_INVALID, "%s: ld (%lld) < n (%d)", name, m.ld, n);
    return GPTPINN_OK;
}
static void add_mat(PackJobs &j, const gptpinn_mat &m, long long nrows, int n, int arr, int seg)
{
    PackMat &pm = j.mat[j.nmat++];
    pm.src = m.ptr; pm.ld = m.ld; pm.nrows = nrows; pm.ncols = n; pm.arr = arr; pm.tile0 = seg;   // seg -> tile0 later
}
static void add_vec(PackJobs &j, const float *src, long long nrows, int vec, int seg)
{
    PackVec &pv = j.vec[j.nvec++];
    pv.src = src; pv.nrows = nrows; pv.vec = vec; pv.tile0 = seg;
}
static int check_common(gptpinn_handle *h, const void *prob, const gptpinn_sweep_args *a, int n, int maxn, int NR, int NI, int NB)
{
    if (!h || !prob || !a) return fail(GPTPINN_E_INVALID, "handle, problem and args are required");
    if (n < 1 || n > maxn) return fail(GPTPINN_E_INVALID, "n = %d outside 1..%d", n, maxn);
    if (NR < 1 || NI < 1 || NB < 1) return fail(GPTPINN_E_INVALID, "NR, NI, NB must be >= 1");
    if (a->Np < 0 || (a->Np > 0 && !a->grid_host)) return fail(GPTPINN_E_INVALID, "grid_host NULL or Np < 0");
    return GPTPINN_OK;
}

extern "C" int gptpinn_kg_sweep(gptpinn_handle *h, const gptpinn_kg_problem *q, const gptpinn_sweep_args *a,
                                gptpinn_sweep_info *info, void *stream)
{
    int rc = check_common(h, q, a, q ? q->n : 0, kMaxN[PDE_KG], q ? q->NR : 0, q ? q->NI : 0, q ? q->NB : 0);
    if (rc) return rc;
    const int n = q->n;
    if ((rc = check_mat(q->P, n, "P")) || (rc = check_mat(q->Pxx, n, "Pxx")) || (rc = check_mat(q->Ptt, n, "Ptt")) ||
        (rc = check_mat(q->PIC, n, "PIC")) || (rc = check_mat(q->PICt, n, "PICt")) || (rc = check_mat(q->PBC, n, "PBC")))
        return rc;
    if (!q->xcos || !q->ICu1 || !q->BCu) return fail(GPTPINN_E_INVALID, "xcos, ICu1 and BCu are required");
    PackJobs jobs;
    memset(&jobs, 0, sizeof(jobs));
    jobs.narr = PdeTraits<PDE_KG>::NARR;
    add_mat(jobs, q->Ptt, q->NR, n, 0, 0); add_mat(jobs, q->Pxx, q->NR, n, 1, 0); add_mat(jobs, q->P, q->NR, n, 2, 0);
    add_vec(jobs, q->xcos, q->NR, 0, 0); add_vec(jobs, q->fhat, q->NR, 1, 0);
    add_mat(jobs, q->PIC, q->NI, n, 0, 1); add_mat(jobs, q->PICt, q->NI, n, 1, 1);
    add_vec(jobs, q->ICu1, q->NI, 0, 1); add_vec(jobs, q->ICu2, q->NI, 1, 1);
    add_mat(jobs, q->PBC, q->NB, n, 0, 2); add_vec(jobs, q->BCu, q->NB, 0, 2);
    Segments seg; seg.nseg = 3; seg.rows[0] = q->NR; seg.rows[1] = q->NI; seg.rows[2] = q->NB; seg.rows[3] = 0;
    std::vector<ParamRow> rows((size_t)a->Np);
    for (int i = 0; i < a->Np; ++i) {
        const double al = a->grid_host[3 * i], be = a->grid_host[3 * i + 1], ga = a->grid_host[3 * i + 2];
        rows[i].tp0 = (float)al; rows[i].tp1 = (float)be;
        rows[i].sp0 = (float)ga; rows[i].sp1 = (float)(2.0 * ga);          // KG_precomp.py:29
        rows[i].idx = i;
    }
    return run_sweep(h, PDE_KG, n, seg, jobs, q->fhat != nullptr, q->NR, q->NI, q->NB, rows, a, info, (cudaStream_t)stream);
}

extern "C" int gptpinn_b_sweep(gptpinn_handle *h, const gptpinn_b_problem *q, const gptpinn_sweep_args *a,
                               gptpinn_sweep_info *info, void *stream)
{
    int rc = check_common(h, q, a, q ? q->n : 0, kMaxN[PDE_B], q ? q->NR : 0, q ? q->NI : 0, q ? q->NB : 0);
    if (rc) return rc;
    const int n = q->n;
    if ((rc = check_mat(q->P, n, "P")) || (rc = check_mat(q->Px, n, "Px")) || (rc = check_mat(q->Pt, n, "Pt")) ||
        (rc = check_mat(q->Pxx, n, "Pxx")) || (rc = check_mat(q->PIC, n, "PIC")) || (rc = check_mat(q->PBC, n, "PBC")))
        return rc;
    if (!q->ICu) return fail(GPTPINN_E_INVALID, "ICu is required");
    PackJobs jobs;
    memset(&jobs, 0, sizeof(jobs));
    jobs.narr = PdeTraits<PDE_B>::NARR;
    add_mat(jobs, q->Pt, q->NR, n, 0, 0); add_mat(jobs, q->Pxx, q->NR, n, 1, 0);
    add_mat(jobs, q->Px, q->NR, n, 2, 0); add_mat(jobs, q->P, q->NR, n, 3, 0);
    add_vec(jobs, q->fhat, q->NR, 0, 0);
    add_mat(jobs, q->PIC, q->NI, n, 0, 1); add_vec(jobs, q->ICu, q->NI, 0, 1);
    add_mat(jobs, q->PBC, q->NB, n, 0, 2); add_vec(jobs, q->BCu, q->NB, 0, 2);
    Segments seg; seg.nseg = 3; seg.rows[0] = q->NR; seg.rows[1] = q->NI; seg.rows[2] = q->NB; seg.rows[3] = 0;
    std::vector<ParamRow> rows((size_t)a->Np);
    for (int i = 0; i < a->Np; ++i) {
        rows[i].tp0 = (float)(-a->grid_host[i]); rows[i].tp1 = 0.f;       // B_precomp.py:19
        rows[i].sp0 = 0.f; rows[i].sp1 = 0.f; rows[i].idx = i;
    }
    return run_sweep(h, PDE_B, n, seg, jobs, q->fhat != nullptr, q->NR, q->NI, q->NB, rows, a, info, (cudaStream_t)stream);
}

extern "C" int gptpinn_ac_sweep(gptpinn_handle *h, const gptpinn_ac_problem *q, const gptpinn_sweep_args *a,
                                gptpinn_sweep_info *info, void *stream)
{
    int rc = check_common(h, q, a, q ? q->n : 0, kMaxN[PDE_AC], q ? q->NR : 0, q ? q->NI : 0, q ? q->NB : 0);
    if (rc) return rc;
    const int n = q->n;
    if ((rc = check_mat(q->P, n, "P")) || (rc = check_mat(q->Pt, n, "Pt")) || (rc = check_mat(q->Pxx, n, "Pxx")) ||
        (rc = check_mat(q->PIC, n, "PIC")) || (rc = check_mat(q->Pub, n, "Pub")) || (rc = check_mat(q->Plb, n, "Plb")) ||
        (rc = check_mat(q->Pdiff, n, "Pdiff")) || (rc = check_mat(q->Pubx, n, "Pubx")) ||
        (rc = check_mat(q->Plbx, n, "Plbx")) || (rc = check_mat(q->Pdiffx, n, "Pdiffx")))
        return rc;
    if (!q->ICu) return fail(GPTPINN_E_INVALID, "ICu is required");
    PackJobs jobs;
    memset(&jobs, 0, sizeof(jobs));
    jobs.narr = PdeTraits<PDE_AC>::NARR;
    add_mat(jobs, q->Pt, q->NR, n, 0, 0); add_mat(jobs, q->Pxx, q->NR, n, 1, 0); add_mat(jobs, q->P, q->NR, n, 2, 0);
    add_vec(jobs, q->fhat, q->NR, 0, 0);
    add_mat(jobs, q->PIC, q->NI, n, 0, 1); add_vec(jobs, q->ICu, q->NI, 0, 1);
    add_mat(jobs, q->Pub, q->NB, n, 0, 2); add_mat(jobs, q->Plb, q->NB, n, 1, 2); add_mat(jobs, q->Pdiff, q->NB, n, 2, 2);
    add_mat(jobs, q->Pubx, q->NB, n, 0, 3); add_mat(jobs, q->Plbx, q->NB, n, 1, 3); add_mat(jobs, q->Pdiffx, q->NB, n, 2, 3);
    Segments seg; seg.nseg = 4; seg.rows[0] = q->NR; seg.rows[1] = q->NI; seg.rows[2] = q->NB; seg.rows[3] = q->NB;
    std::vector<ParamRow> rows((size_t)a->Np);
    for (int i = 0; i < a->Np; ++i) {
        const double lam = a->grid_host[2 * i], eps = a->grid_host[2 * i + 1];
        rows[i].tp0 = (float)(-lam); rows[i].tp1 = (float)(-eps);       // AC_precompute.py:21-22
        if (q->t_given) { rows[i].tp0 = 0.f; rows[i].tp1 = 0.f; }       // (T + 0*Pxx) + 0*P == T bit for bit
        rows[i].sp0 = (float)eps; rows[i].sp1 = (float)(3.0 * eps);     // AC_optimizer.py:42,45
        rows[i].idx = i;
    }
    return run_sweep(h, PDE_AC, n, seg, jobs, q->fhat != nullptr, q->NR, q->NI, q->NB, rows, a, info, (cudaStream_t)stream);
}

extern "C" int gptpinn_argmax_scan(const float *f_dev, const float *m_dev, int Np, float *L_out_dev, int *idx_out_dev, void *stream)
{
    if (!f_dev || !m_dev || !L_out_dev || !idx_out_dev || Np < 0) return fail(GPTPINN_E_INVALID, "gptpinn_argmax_scan: bad arguments");
    CUDA_TRY(launch_argmax_scan(f_dev, m_dev, Np, L_out_dev, idx_out_dev, (cudaStream_t)stream));
    return GPTPINN_OK;
}

extern "C" int gptpinn_mlp_channels(const gptpinn_mlp *mlp, const float *xt_dev, long long N,
                                    float *const out_dev[5], const long long ld[5], void *stream)
{
    if (!mlp || !xt_dev || !out_dev || !ld || N < 0) return fail(GPTPINN_E_INVALID, "gptpinn_mlp_channels: bad arguments");
    if (mlp->nlayers < 2 || mlp->nlayers > MLP_MAX_LAYERS) return fail(GPTPINN_E_INVALID, "nlayers = %d outside 2..%d", mlp->nlayers, MLP_MAX_LAYERS);
    if (mlp->layers[0] != 2 || mlp->layers[mlp->nlayers - 1] != 1) return fail(GPTPINN_E_INVALID, "network must map 2 inputs to 1 output");
    if (mlp->act != GPTPINN_ACT_COS && mlp->act != GPTPINN_ACT_TANH) return fail(GPTPINN_E_INVALID, "unknown activation %d", mlp->act);
    MlpDesc d;
    memset(&d, 0, sizeof(d));
    d.nlayers = mlp->nlayers; d.act = mlp->act;
    for (int l = 0; l < mlp->nlayers; ++l) {
        d.layers[l] = mlp->layers[l];
        if (d.layers[l] < 1 || d.layers[l] > 128) return fail(GPTPINN_E_UNSUPPORTED, "layer width %d outside 1..128", d.layers[l]);
    }
    for (int l = 0; l < mlp->nlayers - 1; ++l) {
        if (!mlp->W[l] || !mlp->b[l]) return fail(GPTPINN_E_INVALID, "W[%d] or b[%d] is NULL", l, l);
        d.W[l] = mlp->W[l]; d.b[l] = mlp->b[l];
    }
    MlpOut o;
    bool any = false;
    for (int c = 0; c < 5; ++c) { o.ptr[c] = out_dev[c]; o.ld[c] = ld[c]; any = any || out_dev[c]; }
    if (!any) return GPTPINN_OK;
    CUDA_TRY(launch_mlp_channels(d, xt_dev, N, o, (cudaStream_t)stream));
    return GPTPINN_OK;
}

// Host-only: the work-item plan the KG sweep would use on a device with `n_sm` SMs (no CUDA call).
extern "C" int gptpinn_kg_plan(const double *grid_host, int Np, int n, int NR, int NI, int NB, int n_sm,
                               const gptpinn_sweep_args *cfg, gptpinn_sweep_info *info, int sl_hist[6], int *covered)
{
    if (!grid_host || !info || !sl_hist || Np < 0 || n < 1 || n > kMaxN[PDE_KG] || n_sm < 1)
        return fail(GPTPINN_E_INVALID, "gptpinn_kg_plan: bad arguments");
    gptpinn_sweep_args a;
    memset(&a, 0, sizeof(a));
    if (cfg) { a.cfg_M = cfg->cfg_M; a.cfg_SL = cfg->cfg_SL; a.cfg_W = cfg->cfg_W; a.cfg_mode = cfg->cfg_mode; a.cfg_K = cfg->cfg_K; }
    std::vector<ParamRow> rows((size_t)Np);
    for (int i = 0; i < Np; ++i) {
        rows[i].tp0 = (float)grid_host[3 * i]; rows[i].tp1 = (float)grid_host[3 * i + 1];
        rows[i].sp0 = (float)grid_host[3 * i + 2]; rows[i].sp1 = (float)(2.0 * grid_host[3 * i + 2]); rows[i].idx = i;
    }
    long long seg[3] = {NR, NI, NB};
    Plan plan;
    int rc = make_plan(rows, PDE_KG, n, n_sm, seg, 3, &a, plan);
    if (rc) return rc;
    memset(info, 0, sizeof(*info));
    info->M = plan.M; info->SL = plan.SL; info->W = plan.W; info->items = (int)plan.items.size(); info->groups = plan.groups;
    info->ctas = plan.ctas; info->rounds = plan.rounds; info->priv = plan.priv; info->K = plan.priv ? plan.K : 1;
    for (int i = 0; i < 6; ++i) sl_hist[i] = 0;
    std::vector<char> seen((size_t)Np, 0);
    int cov = 0;
    for (const ItemDesc &it : plan.items) {
        int l = 0; while ((1 << l) < it.SL) ++l;
        sl_hist[l]++;
        if (it.sib_count > it.SL * plan.M) return fail(GPTPINN_E_INVALID, "plan: item overflows its lanes");
        for (int s = 0; s < it.sib_count; ++s) {
            const int idx = plan.sibs[it.sib_start + s].out_idx;
            if (idx < 0 || idx >= Np || seen[idx]) return fail(GPTPINN_E_INVALID, "plan: parameter %d covered twice or out of range", idx);
            seen[idx] = 1; ++cov;
        }
    }
    if (covered) *covered = cov;
    return GPTPINN_OK;
}

extern "C" int gptpinn_pinn_train_batch(gptpinn_handle *h, int nbatch, const gptpinn_pinn_problem *qs, const gptpinn_mlp *nets,
                                        const gptpinn_train_args *tas, float *status_dev, void *stream)
{
    if (!h || !qs || !nets || !tas || !status_dev) return fail(GPTPINN_E_INVALID, "gptpinn_pinn_train: NULL argument");
    if (nbatch < 1 || nbatch > PINN_MAX_BATCH) return fail(GPTPINN_E_INVALID, "nbatch = %d outside 1..%d", nbatch, PINN_MAX_BATCH);
    const gptpinn_pinn_problem *q = &qs[0];
    const gptpinn_mlp *net = &nets[0];
    const gptpinn_train_args *ta = &tas[0];
    if (q->pde != GPTPINN_PDE_KG && q->pde != GPTPINN_PDE_BURGERS) return fail(GPTPINN_E_INVALID, "pde must be KG (0) or Burgers (1)");
    if (net->nlayers < 3 || net->nlayers > PINN_MAX_HIDDEN + 2) return fail(GPTPINN_E_UNSUPPORTED, "nlayers = %d outside 3..%d", net->nlayers, PINN_MAX_HIDDEN + 2);
    if (net->layers[0] != 2 || net->layers[net->nlayers - 1] != 1) return fail(GPTPINN_E_INVALID, "network must map 2 inputs to 1 output");
    if (q->NR < 1 || q->NI < 1 || q->NB < 1) return fail(GPTPINN_E_INVALID, "NR, NI, NB must be >= 1");
    if (!q->xt_resid || !q->IC_xt || !q->IC_u1 || !q->BC_xt || !q->BC_u) return fail(GPTPINN_E_INVALID, "point sets and targets are required");
    if (q->pde == GPTPINN_PDE_KG && !q->xcos) return fail(GPTPINN_E_INVALID, "xcos is required for Klein-Gordon");
    if (ta->epochs < 0) return fail(GPTPINN_E_INVALID, "epochs < 0");
    for (int t = 1; t < nbatch; ++t) {          // one launch = one network shape, one set of points, one schedule
        const gptpinn_pinn_problem &o = qs[t];
        if (o.pde != q->pde || o.NR != q->NR || o.NI != q->NI || o.NB != q->NB || o.xt_resid != q->xt_resid || o.IC_xt != q->IC_xt ||
            o.BC_xt != q->BC_xt || o.f_hat != q->f_hat || o.xcos != q->xcos || o.IC_u1 != q->IC_u1 || o.IC_u2 != q->IC_u2 || o.BC_u != q->BC_u)
            return fail(GPTPINN_E_INVALID, "batched trainings must share the point sets and targets (training %d differs)", t);
        if (nets[t].nlayers != net->nlayers || nets[t].act != net->act) return fail(GPTPINN_E_INVALID, "batched trainings must share the network shape");
        for (int l = 0; l < net->nlayers; ++l)
            if (nets[t].layers[l] != net->layers[l]) return fail(GPTPINN_E_INVALID, "batched trainings must share the network shape");
        const gptpinn_train_args &u = tas[t];
        if (u.epochs != ta->epochs || u.lr != ta->lr || u.beta1 != ta->beta1 || u.beta2 != ta->beta2 || u.eps != ta->eps || u.tol != ta->tol ||
            u.trace_every != ta->trace_every)
            return fail(GPTPINN_E_INVALID, "batched trainings must share epochs, Adam settings, tolerance and trace stride");
    }
    cudaStream_t st = (cudaStream_t)stream;
    int dev = -1;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev != h->device) return fail(GPTPINN_E_INVALID, "handle belongs to device %d, current device is %d", h->device, dev);
    int coop = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) return fail(GPTPINN_E_UNSUPPORTED, "device does not support cooperative launch");

    PinnDesc d;
    memset(&d, 0, sizeof(d));
    d.nlayers = net->nlayers; d.act = net->act; d.pde = q->pde;
    int P = 0;
    for (int l = 0; l < net->nlayers; ++l) {
        d.layers[l] = net->layers[l];
        if (d.layers[l] < 1 || d.layers[l] > 40) return fail(GPTPINN_E_UNSUPPORTED, "layer width %d outside 1..40", d.layers[l]);
    }
    for (int t = 0; t < nbatch; ++t)
        for (int l = 0; l < net->nlayers - 1; ++l)
            if (!nets[t].W[l] || !nets[t].b[l]) return fail(GPTPINN_E_INVALID, "W[%d] or b[%d] of training %d is NULL", l, l, t);
    for (int l = 0; l < net->nlayers - 1; ++l) P += d.layers[l] * d.layers[l + 1] + d.layers[l + 1];
    d.n_params = P;
    d.NR = q->NR; d.NI = q->NI; d.NB = q->NB;
    d.xt_resid = q->xt_resid; d.IC_xt = q->IC_xt; d.BC_xt = q->BC_xt; d.f_hat = q->f_hat; d.xcos = q->xcos;
    d.IC_u1 = q->IC_u1; d.IC_u2 = q->IC_u2; d.BC_u = q->BC_u;
    d.invR = (float)(1.0 / q->NR); d.invI = (float)(1.0 / q->NI); d.invB = (float)(1.0 / q->NB);
    d.epochs = ta->epochs;
    d.lr = (float)ta->lr; d.beta1 = (float)ta->beta1; d.beta2 = (float)ta->beta2;
    d.lr_d = ta->lr; d.beta1_d = ta->beta1; d.beta2_d = ta->beta2;
    d.omb1 = (float)(1.0 - ta->beta1); d.omb2 = (float)(1.0 - ta->beta2); d.eps = (float)ta->eps;
    d.tol = ta->tol >= 0 ? (float)ta->tol : -1.f;
    d.trace_every = ta->trace_every;

    int blocks = 0, group = 0;
    size_t smem = 0;
    CUDA_TRY(plan_pinn_train(d, h->n_sm, nbatch, &blocks, &smem, &group));
    const int stride = (P + 4 + 31) & ~31;
    const int P4 = (P + 3) & ~3;
    const size_t per = (size_t)3 * P4 + (size_t)group * stride;                       // floats per training
    const size_t need = (per * nbatch + 64) * sizeof(float);
    int rc = ensure(&h->train_ws, &h->train_cap, need);
    if (rc) return rc;
    float *ws = (float *)h->train_ws;
    d.n_partials = group; d.partial_stride = stride;
    float *tail = ws + per * nbatch;
    d.stop = (int *)tail;
    PinnBatch bt;
    memset(&bt, 0, sizeof(bt));
    bt.n = nbatch; bt.group = group;
    CUDA_TRY(cudaMemsetAsync(tail, 0, 64 * sizeof(float), st));
    CUDA_TRY(cudaMemsetAsync(status_dev, 0, (size_t)2 * nbatch * sizeof(float), st));
    for (int t = 0; t < nbatch; ++t) {
        PinnTrainee &T = bt.t[t];
        float *base = ws + per * t;
        T.params = base; T.adam_m = base + P4; T.adam_v = base + 2 * P4; T.partials = base + 3 * P4;
        T.status = status_dev + 2 * t;
        T.grad_out = tas[t].grad_out_dev;
        T.loss_trace = ta->trace_every > 0 ? tas[t].loss_trace_dev : nullptr;
        if (q->pde == GPTPINN_PDE_KG) { T.pa = (float)qs[t].p0; T.pb = (float)qs[t].p1; T.pc = (float)qs[t].p2; }
        else { T.pa = (float)(-qs[t].p0); T.pb = 0.f; T.pc = 0.f; }
        CUDA_TRY(cudaMemsetAsync(T.adam_m, 0, (size_t)2 * P4 * sizeof(float), st));      // fresh optimizer state
        int off = 0;
        for (int l = 0; l < net->nlayers - 1; ++l) {                                     // pack: W then b, layer after layer
            const int nw = d.layers[l] * d.layers[l + 1], nb = d.layers[l + 1];
            CUDA_TRY(cudaMemcpyAsync(base + off, nets[t].W[l], nw * sizeof(float), cudaMemcpyDeviceToDevice, st)); off += nw;
            CUDA_TRY(cudaMemcpyAsync(base + off, nets[t].b[l], nb * sizeof(float), cudaMemcpyDeviceToDevice, st)); off += nb;
        }
    }
    if (ta->epochs > 0) CUDA_TRY(launch_pinn_train(d, bt, blocks, smem, st));
    for (int t = 0; t < nbatch; ++t) {
        const float *base = ws + per * t;
        int off = 0;
        for (int l = 0; l < net->nlayers - 1; ++l) {                                     // unpack into the caller's tensors
            const int nw = d.layers[l] * d.layers[l + 1], nb = d.layers[l + 1];
            CUDA_TRY(cudaMemcpyAsync((void *)nets[t].W[l], base + off, nw * sizeof(float), cudaMemcpyDeviceToDevice, st)); off += nw;
            CUDA_TRY(cudaMemcpyAsync((void *)nets[t].b[l], base + off, nb * sizeof(float), cudaMemcpyDeviceToDevice, st)); off += nb;
        }
    }
    return GPTPINN_OK;
}

extern "C" int gptpinn_pinn_train(gptpinn_handle *h, const gptpinn_pinn_problem *q, const gptpinn_mlp *net,
                                  const gptpinn_train_args *ta, float *status_dev, void *stream)
{
    if (!q || !net || !ta) return fail(GPTPINN_E_INVALID, "gptpinn_pinn_train: NULL argument");
    return gptpinn_pinn_train_batch(h, 1, q, net, ta, status_dev, stream);
}
