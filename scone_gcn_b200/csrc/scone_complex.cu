// scone_complex.cu — host-side exact integer construction of the device-resident complex.
//
// Replaces the dense operator assembly of the reference
//   L1_lower = B1.T @ B1, L1_upper = B2 @ B2.T, ebli: L1 = L_lower + L_upper, L1 @ L1
//     (trajectory_analysis/trajectory_experiments.py:239-253)
//   nbrhoods / B1_jax / Bconds_func (trajectory_experiments.py:272-303)
// with CSR index arrays built from the signed incidence lists (2 nonzeros per B1 column, 3 per B2
// column).  All arithmetic here is int32 and exact; coefficients are stored as fp32 (|v| <= 2^24).
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include "common.cuh"

static thread_local char g_err[1024] = "";
std::atomic<long long> g_scone_launches{0};

void scone_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* scone_last_error(void) { return g_err; }
extern "C" int scone_version(void) { return SCONE_B200_VERSION; }
extern "C" int64_t scone_launch_count(void) { return g_scone_launches.load(); }

// ---- optional per-kernel timing ---------------------------------------------------------------
bool g_scone_prof = false;
namespace {
struct ProfRec { int kind; cudaEvent_t a, b; };
std::vector<ProfRec> g_prof_pending;
std::vector<cudaEvent_t> g_prof_pool;
double g_prof_ms[SCONE_K_COUNT] = {0};
long long g_prof_n[SCONE_K_COUNT] = {0};
cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
void prof_collect() {
    for (auto& r : g_prof_pending) {
        cudaEventSynchronize(r.b);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { g_prof_ms[r.kind] += ms; g_prof_n[r.kind] += 1; }
        g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b);
    }
    g_prof_pending.clear();
}
}  // namespace
void scone_prof_begin_impl(int kind, cudaStream_t st) {
    ProfRec r{kind, prof_event(), prof_event()};
    cudaEventRecord(r.a, st);
    g_prof_pending.push_back(r);
}
void scone_prof_end_impl(int kind, cudaStream_t st) {
    for (size_t i = g_prof_pending.size(); i-- > 0;)
        if (g_prof_pending[i].kind == kind) { cudaEventRecord(g_prof_pending[i].b, st); break; }
}
static unsigned long long* g_prof_rows_dev = nullptr;      // [SCONE_K_COUNT]
unsigned long long* scone_prof_row_counter(int kind) {
    return (g_scone_prof && g_prof_rows_dev) ? g_prof_rows_dev + kind : nullptr;
}
extern "C" int scone_profile_enable(int32_t on) {
    prof_collect();
    if (on && !g_prof_rows_dev) {
        SCONE_CUDA(cudaMalloc((void**)&g_prof_rows_dev, SCONE_K_COUNT * sizeof(unsigned long long)));
        SCONE_CUDA(cudaMemset(g_prof_rows_dev, 0, SCONE_K_COUNT * sizeof(unsigned long long)));
    }
    g_scone_prof = on != 0;
    return 0;
}
extern "C" int scone_profile_reset(void) {
    prof_collect();
    for (int k = 0; k < SCONE_K_COUNT; ++k) { g_prof_ms[k] = 0; g_prof_n[k] = 0; }
    if (g_prof_rows_dev) SCONE_CUDA(cudaMemset(g_prof_rows_dev, 0, SCONE_K_COUNT * sizeof(unsigned long long)));
    return 0;
}
/* rows produced by the flagged unit kernels of `kind` (0 fwd, 1 bwd) since the last reset; synchronises the device */
extern "C" int scone_profile_read_rows(int32_t kind, int64_t* rows) {
    SCONE_REQUIRE(kind >= 0 && kind < SCONE_K_COUNT && rows, "scone_profile_read_rows: bad arguments");
    *rows = 0;
    if (!g_prof_rows_dev) return 0;
    unsigned long long v = 0;
    SCONE_CUDA(cudaDeviceSynchronize());
    SCONE_CUDA(cudaMemcpy(&v, g_prof_rows_dev + kind, sizeof(v), cudaMemcpyDeviceToHost));
    *rows = (int64_t)v;
    return 0;
}
extern "C" int scone_profile_read(int32_t kind, int64_t* launches, double* total_ms) {
    SCONE_REQUIRE(kind >= 0 && kind < SCONE_K_COUNT, "scone_profile_read: kind out of range");
    prof_collect();
    if (launches) *launches = g_prof_n[kind];
    if (total_ms) *total_ms = g_prof_ms[kind];
    return 0;
}

namespace {

struct IntCsr {
    std::vector<int32_t> rowptr, col, val;
};

// C = A * B for square int CSR matrices with sorted columns; drops exact zeros.
IntCsr spgemm(const IntCsr& A, const IntCsr& B, int32_t n) {
    IntCsr C;
    C.rowptr.assign(n + 1, 0);
    std::vector<int32_t> acc(n, 0), mark(n, -1), touched;
    for (int32_t i = 0; i < n; ++i) {
        touched.clear();
        for (int32_t p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) {
            int32_t k = A.col[p], a = A.val[p];
            for (int32_t q = B.rowptr[k]; q < B.rowptr[k + 1]; ++q) {
                int32_t j = B.col[q];
                if (mark[j] != i) {
                    mark[j] = i;
                    acc[j] = 0;
                    touched.push_back(j);
                }
                acc[j] += a * B.val[q];
            }
        }
        std::sort(touched.begin(), touched.end());
        for (int32_t j : touched)
            if (acc[j] != 0) {
                C.col.push_back(j);
                C.val.push_back(acc[j]);
            }
        C.rowptr[i + 1] = (int32_t)C.col.size();
    }
    return C;
}

// C = A + B (sorted merge), drops exact zeros (the shared-triangle off-diagonals of L1 cancel).
IntCsr spadd(const IntCsr& A, const IntCsr& B, int32_t n) {
    IntCsr C;
    C.rowptr.assign(n + 1, 0);
    for (int32_t i = 0; i < n; ++i) {
        int32_t p = A.rowptr[i], pe = A.rowptr[i + 1], q = B.rowptr[i], qe = B.rowptr[i + 1];
        while (p < pe || q < qe) {
            int32_t c, v;
            if (q >= qe || (p < pe && A.col[p] < B.col[q])) { c = A.col[p]; v = A.val[p]; ++p; }
            else if (p >= pe || B.col[q] < A.col[p]) { c = B.col[q]; v = B.val[q]; ++q; }
            else { c = A.col[p]; v = A.val[p] + B.val[q]; ++p; ++q; }
            if (v != 0) { C.col.push_back(c); C.val.push_back(v); }
        }
        C.rowptr[i + 1] = (int32_t)C.col.size();
    }
    return C;
}

// L = M^T M restricted to the structure "rows of M are simplices of the other dimension":
// given for every column-simplex (edge) its incident row-simplices with signs, and for every
// row-simplex its incident edges with signs, L[e][e'] = sum_r s(r,e) s(r,e').
IntCsr gram(int32_t n_edges, const std::vector<std::vector<std::pair<int32_t, int32_t>>>& edge_to_rows,
            const std::vector<std::vector<std::pair<int32_t, int32_t>>>& row_to_edges) {
    IntCsr L;
    L.rowptr.assign(n_edges + 1, 0);
    std::map<int32_t, int32_t> acc;
    for (int32_t e = 0; e < n_edges; ++e) {
        acc.clear();
        for (auto& rs : edge_to_rows[e])
            for (auto& es : row_to_edges[rs.first]) acc[es.first] += rs.second * es.second;
        for (auto& kv : acc)
            if (kv.second != 0) { L.col.push_back(kv.first); L.val.push_back(kv.second); }
        L.rowptr[e + 1] = (int32_t)L.col.size();
    }
    return L;
}

// ---- internal edge order -------------------------------------------------------------------------------------------
// Activations are private to the library, so edge rows may live in any order on the device.  The reference's order
// (lexicographic over nodes sorted by x+y) spreads a spatial neighbourhood over the whole index range: a window of 256
// consecutive edges touches ~1170 distinct neighbour rows on the 100k-node complex.  We order the NODES by recursive
// breadth-first bisection of the graph (no coordinates needed): inside a part, breadth-first search from a
// pseudo-peripheral node, the first half of the visit order goes left, the second half right, recurse; edges follow their
// earlier end node.  Every index window is then a compact patch of the mesh at every scale (256 edges -> ~450 distinct
// neighbour rows, within 15 % of a Hilbert curve over the true coordinates), which is what lets the gathers of
// consecutive edges hit in L1 / L2.  (BFS-distance embeddings are not usable here: the long convex-hull edges of a
// Delaunay complex are hop-distance shortcuts around the whole boundary.)
struct BfsScratch {
    std::vector<int32_t> stamp, queue;
    int32_t cur = 0;
};

// BFS inside the part perm[lo, hi) (membership: lo <= pos[v] < hi) from src; visit order in sc.queue; returns #visited
int32_t bfs_part(const std::vector<int32_t>& ptr, const std::vector<int32_t>& adj, const std::vector<int32_t>& pos, int32_t lo,
                 int32_t hi, int32_t src, BfsScratch& sc) {
    ++sc.cur;
    sc.queue.clear();
    sc.queue.push_back(src);
    sc.stamp[src] = sc.cur;
    for (size_t h = 0; h < sc.queue.size(); ++h) {
        const int32_t u = sc.queue[h];
        for (int32_t p = ptr[u]; p < ptr[u + 1]; ++p) {
            const int32_t v = adj[p];
            if (sc.stamp[v] != sc.cur && pos[v] >= lo && pos[v] < hi) {
                sc.stamp[v] = sc.cur;
                sc.queue.push_back(v);
            }
        }
    }
    return (int32_t)sc.queue.size();
}

// order[new] = old edge id
std::vector<int32_t> locality_order(int32_t N, int32_t E, const int32_t* edge_nodes) {
    std::vector<int32_t> ptr(N + 1, 0), adj(2 * (size_t)E);
    for (int32_t e = 0; e < E; ++e) { ptr[edge_nodes[2 * e] + 1]++; ptr[edge_nodes[2 * e + 1] + 1]++; }
    for (int32_t n = 0; n < N; ++n) ptr[n + 1] += ptr[n];
    std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);
    for (int32_t e = 0; e < E; ++e) {
        int32_t a = edge_nodes[2 * e], b = edge_nodes[2 * e + 1];
        adj[fill[a]++] = b;
        adj[fill[b]++] = a;
    }
    // perm: connected nodes first (isolated nodes carry no edge rows), pos = inverse
    std::vector<int32_t> perm, pos(N);
    perm.reserve(N);
    for (int32_t n = 0; n < N; ++n) if (ptr[n + 1] > ptr[n]) perm.push_back(n);
    const int32_t n_conn = (int32_t)perm.size();
    for (int32_t n = 0; n < N; ++n) if (ptr[n + 1] == ptr[n]) perm.push_back(n);
    for (int32_t i = 0; i < N; ++i) pos[perm[i]] = i;
    BfsScratch sc;
    sc.stamp.assign(N, 0);
    constexpr int32_t kLeaf = 16;
    std::vector<std::pair<int32_t, int32_t>> stack;
    stack.push_back({0, n_conn});
    std::vector<int32_t> rest;
    while (!stack.empty()) {
        const int32_t lo = stack.back().first, hi = stack.back().second;
        stack.pop_back();
        if (hi - lo <= kLeaf) continue;
        int32_t cnt = bfs_part(ptr, adj, pos, lo, hi, perm[lo], sc);
        if (cnt < hi - lo) {                         // disconnected part: [component of perm[lo]] [everything else]
            rest.clear();
            for (int32_t i = lo; i < hi; ++i) if (sc.stamp[perm[i]] != sc.cur) rest.push_back(perm[i]);
            for (int32_t i = 0; i < cnt; ++i) perm[lo + i] = sc.queue[i];
            for (size_t i = 0; i < rest.size(); ++i) perm[lo + cnt + (int32_t)i] = rest[i];
            for (int32_t i = lo; i < hi; ++i) pos[perm[i]] = i;
            stack.push_back({lo + cnt, hi});
            stack.push_back({lo, lo + cnt});
            continue;
        }
        const int32_t far = sc.queue.back();         // pseudo-peripheral node of the part
        cnt = bfs_part(ptr, adj, pos, lo, hi, far, sc);
        for (int32_t i = 0; i < cnt; ++i) perm[lo + i] = sc.queue[i];
        for (int32_t i = lo; i < hi; ++i) pos[perm[i]] = i;
        const int32_t mid = lo + (hi - lo) / 2;
        stack.push_back({mid, hi});
        stack.push_back({lo, mid});
    }
    std::vector<std::pair<uint64_t, int32_t>> key(E);
    for (int32_t e = 0; e < E; ++e) {
        const uint32_t pa = (uint32_t)pos[edge_nodes[2 * e]], pb = (uint32_t)pos[edge_nodes[2 * e + 1]];
        key[e] = {((uint64_t)std::min(pa, pb) << 32) | std::max(pa, pb), e};
    }
    std::sort(key.begin(), key.end());
    std::vector<int32_t> order(E);
    for (int32_t i = 0; i < E; ++i) order[i] = key[i].second;
    return order;
}

template <typename T>
int upload(T** dst, const std::vector<T>& src) {
    size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    SCONE_CUDA(cudaMalloc((void**)dst, bytes));
    if (!src.empty()) SCONE_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

}  // namespace

static int complex_create_impl(int32_t N, int32_t E, int32_t F, const int32_t* edge_nodes, const int8_t* edge_signs,
                               const int32_t* tri_edges, const int8_t* tri_signs, int32_t model, bool upload_to_device,
                               scone_complex** out) {
    SCONE_REQUIRE(out != nullptr, "scone_complex_create: out is NULL");
    *out = nullptr;
    SCONE_REQUIRE(N > 0 && E > 0 && F >= 0, "scone_complex_create: bad sizes N=%d E=%d F=%d", N, E, F);
    SCONE_REQUIRE(edge_nodes != nullptr && (F == 0 || (tri_edges != nullptr && tri_signs != nullptr)),
                  "scone_complex_create: NULL index array");
    SCONE_REQUIRE(model == SCONE_MODEL_SCONE || model == SCONE_MODEL_EBLI,
                  "scone_complex_create: model must be scone (0) or ebli (1), got %d", model);
    using PairList = std::vector<std::vector<std::pair<int32_t, int32_t>>>;
    PairList edge_to_nodes(E), node_to_edges(N), edge_to_tris(E), tri_to_edges(F);
    for (int32_t e = 0; e < E; ++e) {
        for (int k = 0; k < 2; ++k) {
            int32_t n = edge_nodes[2 * e + k];
            SCONE_REQUIRE(n >= 0 && n < N, "scone_complex_create: edge %d has node %d outside [0,%d)", e, n, N);
            int32_t s = edge_signs ? edge_signs[2 * e + k] : (k == 0 ? -1 : 1);
            SCONE_REQUIRE(s == 1 || s == -1, "scone_complex_create: edge %d sign %d not +-1", e, s);
            edge_to_nodes[e].push_back({n, s});
            node_to_edges[n].push_back({e, s});          // e ascending by construction
        }
        SCONE_REQUIRE(edge_nodes[2 * e] != edge_nodes[2 * e + 1], "scone_complex_create: edge %d is a self loop", e);
    }
    for (int32_t f = 0; f < F; ++f)
        for (int k = 0; k < 3; ++k) {
            int32_t e = tri_edges[3 * f + k], s = tri_signs[3 * f + k];
            SCONE_REQUIRE(e >= 0 && e < E, "scone_complex_create: triangle %d has edge %d outside [0,%d)", f, e, E);
            SCONE_REQUIRE(s == 1 || s == -1, "scone_complex_create: triangle %d sign %d not +-1", f, s);
            edge_to_tris[e].push_back({f, s});
            tri_to_edges[f].push_back({e, s});
        }

    IntCsr Ll = gram(E, edge_to_nodes, node_to_edges);     // B1^T B1
    IntCsr Lu = gram(E, edge_to_tris, tri_to_edges);       // B2 B2^T
    IntCsr S0, S1;
    if (model == SCONE_MODEL_SCONE) {
        S0 = std::move(Ll);
        S1 = std::move(Lu);
    } else {
        S0 = spadd(Ll, Lu, E);
        S1 = spgemm(S0, S0, E);
    }

    scone_complex* cx = new scone_complex();
    cx->N = N; cx->E = E; cx->F = F; cx->model = model;
    IntCsr* src[2] = {&S0, &S1};
    for (int k = 0; k < 2; ++k) {
        cx->hS[k].rowptr = src[k]->rowptr;
        cx->hS[k].col = src[k]->col;
        cx->hS[k].val.resize(src[k]->val.size());
        for (size_t i = 0; i < src[k]->val.size(); ++i) {
            SCONE_REQUIRE(std::abs(src[k]->val[i]) < (1 << 24), "shift coefficient not exactly representable in fp32");
            cx->hS[k].val[i] = (float)src[k]->val[i];
        }
    }
    // neighbour table: sorted ascending, padded with -1 (trajectory_experiments.py:279)
    std::vector<std::vector<int32_t>> nbrs(N);
    for (int32_t e = 0; e < E; ++e) {
        nbrs[edge_nodes[2 * e]].push_back(edge_nodes[2 * e + 1]);
        nbrs[edge_nodes[2 * e + 1]].push_back(edge_nodes[2 * e]);
    }
    int32_t D = 0;
    for (auto& v : nbrs) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        D = std::max<int32_t>(D, (int32_t)v.size());
    }
    cx->D = D;
    cx->h_nbrhoods.assign((size_t)N * std::max(D, 1), -1);
    for (int32_t n = 0; n < N; ++n)
        for (size_t j = 0; j < nbrs[n].size(); ++j) cx->h_nbrhoods[(size_t)n * D + j] = nbrs[n][j];

    // device arrays live in the internal (locality) edge order: rank[old] = new, order[new] = old
    std::vector<int32_t> order(E), rank(E);
    const char* no_reorder = getenv("SCONE_B200_NO_REORDER");
    if (no_reorder && no_reorder[0] == '1') {
        for (int32_t e = 0; e < E; ++e) order[e] = e;
    } else {
        order = locality_order(N, E, edge_nodes);
    }
    for (int32_t i = 0; i < E; ++i) rank[order[i]] = i;
    cx->h_rank = rank;
    if (!upload_to_device) {                 // index-only handle: getters work, every device op is refused
        cx->host_only = true;
        *out = cx;
        return 0;
    }
    int dev = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) {
        scone_set_error("scone_complex_create: no CUDA device (%s); this library has no CPU path", cudaGetErrorString(ce));
        delete cx;
        return 3;
    }
    cudaDeviceGetAttribute(&cx->num_sms, cudaDevAttrMultiProcessorCount, dev);

    int rc = 0;
    for (int k = 0; k < 2 && !rc; ++k) {
        const HostCsr& h = cx->hS[k];
        std::vector<int32_t> rowptr(E + 1, 0);
        std::vector<int2> ent(h.col.size());
        std::vector<std::pair<int32_t, int32_t>> row;
        size_t w = 0;
        for (int32_t i = 0; i < E; ++i) {
            const int32_t o = order[i];
            row.clear();
            for (int32_t p = h.rowptr[o]; p < h.rowptr[o + 1]; ++p) {
                int32_t bits;
                memcpy(&bits, &h.val[p], 4);
                row.push_back({rank[h.col[p]], bits});
            }
            std::sort(row.begin(), row.end());           // ascending internal column: fixed summation order
            for (auto& cv : row) ent[w++] = make_int2(cv.first, cv.second);
            rowptr[i + 1] = (int32_t)w;
        }
        rc |= upload(&cx->d_rowptr[k], rowptr);
        rc |= upload(&cx->d_ent[k], ent);
    }
    {   // merged rows: one pass over a neighbour row feeds the own-row term and both operator sums
        std::vector<int32_t> mptr(E + 1, 0);
        std::vector<int2> ment;
        ment.reserve(cx->hS[0].col.size() + (size_t)E);
        std::map<int32_t, std::pair<int32_t, int32_t>> mrow;
        bool fits = true;
        for (int32_t i = 0; i < E; ++i) {
            const int32_t o = order[i];
            mrow.clear();
            mrow[i] = {0, 0};                            // the diagonal is always present (own-row term)
            for (int k = 0; k < 2; ++k) {
                const HostCsr& h = cx->hS[k];
                for (int32_t p = h.rowptr[o]; p < h.rowptr[o + 1]; ++p) {
                    const int32_t v = (int32_t)h.val[p];
                    if (v < -32768 || v > 32767) fits = false;
                    auto& cc = mrow[rank[h.col[p]]];
                    (k == 0 ? cc.first : cc.second) = v;
                }
            }
            for (auto& kv : mrow) {
                const uint32_t pk = ((uint32_t)(uint16_t)(int16_t)kv.second.second << 16) | (uint32_t)(uint16_t)(int16_t)kv.second.first;
                ment.push_back(make_int2(kv.first, (int32_t)pk));
            }
            mptr[i + 1] = (int32_t)ment.size();
        }
        if (fits && !rc) {                               // coefficients beyond int16: the slab kernels are simply not used
            rc |= upload(&cx->d_mptr, mptr);
            rc |= upload(&cx->d_ment, ment);
        }
    }
    std::vector<int32_t> inc_ptr(N + 1, 0);
    std::vector<int2> inc_ent;
    inc_ent.reserve(2 * (size_t)E);
    std::vector<std::pair<int32_t, int32_t>> row;
    for (int32_t n = 0; n < N; ++n) {
        row.clear();
        for (auto& es : node_to_edges[n]) {
            float sgn = (float)es.second;
            int32_t bits;
            memcpy(&bits, &sgn, 4);
            row.push_back({rank[es.first], bits});
        }
        std::sort(row.begin(), row.end());
        for (auto& cv : row) inc_ent.push_back(make_int2(cv.first, cv.second));
        inc_ptr[n + 1] = (int32_t)inc_ent.size();
    }
    if (!rc) rc |= upload(&cx->d_inc_ptr, inc_ptr);
    if (!rc) rc |= upload(&cx->d_inc_ent, inc_ent);
    if (!rc) rc |= upload(&cx->d_rank, rank);
    if (!rc) rc |= upload(&cx->d_nbrhoods, cx->h_nbrhoods);
    if (rc) {
        scone_complex_destroy(cx);
        return rc;
    }
    *out = cx;
    return 0;
}

extern "C" int scone_complex_create(int32_t N, int32_t E, int32_t F, const int32_t* edge_nodes, const int8_t* edge_signs,
                                    const int32_t* tri_edges, const int8_t* tri_signs, int32_t model, scone_complex** out) {
    return complex_create_impl(N, E, F, edge_nodes, edge_signs, tri_edges, tri_signs, model, true, out);
}

extern "C" int scone_complex_create_index_only(int32_t N, int32_t E, int32_t F, const int32_t* edge_nodes,
                                               const int8_t* edge_signs, const int32_t* tri_edges, const int8_t* tri_signs,
                                               int32_t model, scone_complex** out) {
    return complex_create_impl(N, E, F, edge_nodes, edge_signs, tri_edges, tri_signs, model, false, out);
}

extern "C" int scone_complex_destroy(scone_complex* cx) {
    if (!cx) return 0;
    for (int k = 0; k < 2; ++k) {
        cudaFree(cx->d_rowptr[k]);
        cudaFree(cx->d_ent[k]);
    }
    cudaFree(cx->d_nbrhoods);
    cudaFree(cx->d_inc_ptr);
    cudaFree(cx->d_inc_ent);
    cudaFree(cx->d_rank);
    cudaFree(cx->d_mptr);
    cudaFree(cx->d_ment);
    delete cx;
    return 0;
}

extern "C" int scone_complex_dims(const scone_complex* cx, int32_t* N, int32_t* E, int32_t* F, int32_t* D, int64_t* nnz0,
                                  int64_t* nnz1) {
    SCONE_REQUIRE(cx != nullptr, "scone_complex_dims: NULL complex");
    if (N) *N = cx->N;
    if (E) *E = cx->E;
    if (F) *F = cx->F;
    if (D) *D = cx->D;
    if (nnz0) *nnz0 = (int64_t)cx->hS[0].col.size();
    if (nnz1) *nnz1 = (int64_t)cx->hS[1].col.size();
    return 0;
}

extern "C" int scone_complex_get_shift_csr(const scone_complex* cx, int32_t which, int32_t* rowptr, int32_t* col,
                                           float* val) {
    SCONE_REQUIRE(cx != nullptr && (which == 0 || which == 1), "scone_complex_get_shift_csr: bad arguments");
    const HostCsr& h = cx->hS[which];
    if (rowptr) memcpy(rowptr, h.rowptr.data(), h.rowptr.size() * sizeof(int32_t));
    if (col) memcpy(col, h.col.data(), h.col.size() * sizeof(int32_t));
    if (val) memcpy(val, h.val.data(), h.val.size() * sizeof(float));
    return 0;
}

extern "C" int scone_complex_get_edge_rank(const scone_complex* cx, int32_t* rank) {
    SCONE_REQUIRE(cx != nullptr && rank != nullptr, "scone_complex_get_edge_rank: bad arguments");
    memcpy(rank, cx->h_rank.data(), cx->h_rank.size() * sizeof(int32_t));
    return 0;
}

extern "C" int scone_complex_get_nbrhoods(const scone_complex* cx, int32_t* nbrhoods) {
    SCONE_REQUIRE(cx != nullptr && nbrhoods != nullptr, "scone_complex_get_nbrhoods: bad arguments");
    memcpy(nbrhoods, cx->h_nbrhoods.data(), cx->h_nbrhoods.size() * sizeof(int32_t));
    return 0;
}
