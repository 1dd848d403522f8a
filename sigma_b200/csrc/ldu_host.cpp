// ldu_host.cpp -- host-side index work of the ILDU(0) preconditioner (SURVEY.md
// 8f rank 4): the sparsity patterns of the factors and the level schedules
// that let the device run the factorisation and the two triangular solves row-
// parallel.  Pure int32 work with no CUDA call, so it runs (and is tested)
// without a GPU; the patterns must equal the oracle's restatement of
// incomplete_ldu_sparsity_pattern (src/solver/ldu_solvers.f90:396-441) bit for
// bit.
//
// The reference is serial: row i of the factorisation (:331-381) reads rows k
// of U and D(k) for the lower neighbours k of i, and lower_triangular_solve /
// upper_triangular_solve (:208-263) read x(j) for the stored neighbours j of
// row i.  Those are the only dependencies; rows whose dependencies are all
// satisfied form a level and can run concurrently while every row still does
// its own arithmetic in the reference's order.
#include <algorithm>
#include <vector>

#include "internal.h"

using namespace sigb;

extern "C" {

// Input: the matrix as compressed ROWS in the order its entry iterator produces
// them (a csr_matrix's stored arrays; for csc / ellpack sources the row form of
// sigb_matrix_copy, which keeps iteration order), 1-based.
//
// Output (all sized by the caller: ptr arrays n + 1, node arrays ne, dest ne,
// row lists n, level pointers n + 1):
//   Lptr/Lnode, Uptr/Unode : strict lower / upper patterns, each row in the
//                            source's order (ll_graph add_edge order, :424-432)
//   dest[e]                : where entry e of the source goes in the combined
//                            value array [ Lval | Uval | D ] (0-based offset):
//                            "Copy A into L, D, U", :307-324
//   frows, flev            : rows (1-based) grouped by factorisation / forward-
//                            solve level, ascending inside a level; level l is
//                            frows[flev[l] .. flev[l+1]), *nflev levels
//   brows, blev            : the same for the backward solve with U
int sigb_ldu_symbolic(int32_t n, const int32_t *ptr1, const int32_t *node1, int32_t *Lptr, int32_t *Lnode,
                      int32_t *Uptr, int32_t *Unode, int64_t *dest, int32_t *frows, int32_t *flev, int32_t *nflev,
                      int32_t *brows, int32_t *blev, int32_t *nblev)
{
    SIGB_REQUIRE(n >= 0 && ptr1 && Lptr && Uptr && dest && frows && flev && nflev && brows && blev && nblev,
                 SIGB_ERR_ARG, "sigb_ldu_symbolic: bad argument");
    const int64_t ne = (int64_t)ptr1[n] - 1;
    SIGB_REQUIRE(ne == 0 || (node1 && Lnode && Unode), SIGB_ERR_ARG, "sigb_ldu_symbolic: bad argument");
    // patterns: row i keeps its lower / upper neighbours in stored order
    Lptr[0] = Uptr[0] = 1;
    for (int32_t i = 0; i < n; i++) {
        int32_t nl = 0, nu = 0;
        for (int32_t k = ptr1[i] - 1; k < ptr1[i + 1] - 1; k++) {
            const int32_t j = node1[k];
            SIGB_REQUIRE(j >= 1 && j <= n, SIGB_ERR_ARG, "sigb_ldu_symbolic: column %d outside 1..%d", j, n);
            if (j < i + 1) Lnode[Lptr[i] - 1 + nl++] = j;
            else if (j > i + 1) Unode[Uptr[i] - 1 + nu++] = j;
        }
        Lptr[i + 1] = Lptr[i] + nl;
        Uptr[i + 1] = Uptr[i] + nu;
    }
    const int64_t nL = (int64_t)Lptr[n] - 1, nU = (int64_t)Uptr[n] - 1;
    // destinations in [ Lval | Uval | D ]
    for (int32_t i = 0; i < n; i++) {
        int32_t nl = 0, nu = 0;
        for (int32_t k = ptr1[i] - 1; k < ptr1[i + 1] - 1; k++) {
            const int32_t j = node1[k];
            if (j < i + 1) dest[k] = (int64_t)Lptr[i] - 1 + nl++;
            else if (j > i + 1) dest[k] = nL + (int64_t)Uptr[i] - 1 + nu++;
            else dest[k] = nL + nU + i;
        }
    }
    // forward levels: level(i) = 1 + max level of its lower neighbours
    std::vector<int32_t> lev((size_t)n, 0);
    int32_t maxl = n > 0 ? 0 : -1;
    for (int32_t i = 0; i < n; i++) {
        int32_t l = 0;
        for (int32_t k = Lptr[i] - 1; k < Lptr[i + 1] - 1; k++) l = std::max(l, lev[(size_t)Lnode[k] - 1] + 1);
        lev[(size_t)i] = l;
        maxl = std::max(maxl, l);
    }
    auto bucket = [&](int32_t nlev, int32_t *rows, int32_t *lptr) {
        std::vector<int32_t> cnt((size_t)nlev + 1, 0);
        for (int32_t i = 0; i < n; i++) cnt[(size_t)lev[(size_t)i] + 1]++;
        lptr[0] = 0;
        for (int32_t l = 0; l < nlev; l++) lptr[l + 1] = lptr[l] + cnt[(size_t)l + 1];
        std::vector<int32_t> fill(lptr, lptr + nlev);
        for (int32_t i = 0; i < n; i++) rows[fill[(size_t)lev[(size_t)i]]++] = i + 1;   // ascending inside a level
    };
    *nflev = maxl + 1;
    bucket(*nflev, frows, flev);
    // backward levels: level(i) = 1 + max level of its upper neighbours, rows descending
    std::fill(lev.begin(), lev.end(), 0);
    maxl = n > 0 ? 0 : -1;
    for (int32_t i = n - 1; i >= 0; i--) {
        int32_t l = 0;
        for (int32_t k = Uptr[i] - 1; k < Uptr[i + 1] - 1; k++) l = std::max(l, lev[(size_t)Unode[k] - 1] + 1);
        lev[(size_t)i] = l;
        maxl = std::max(maxl, l);
    }
    *nblev = maxl + 1;
    bucket(*nblev, brows, blev);
    return SIGB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Statically scheduled sweeps (ldu_sweep.h): the plan
// ---------------------------------------------------------------------------
#include "ldu_sweep.h"

namespace sigb {

void build_sweep_plan(int32_t n, const int32_t *ptr1, const int32_t *node1, int backward, int64_t levels,
                      SweepPlan &P, bool fill_slots)
{
    P = SweepPlan();
    P.n = n;
    P.backward = backward;
    if (n <= 0) return;
    const int64_t ne = (int64_t)ptr1[n] - 1;
    auto pos_of = [&](int32_t row1) -> int64_t { return backward ? (int64_t)n - row1 : (int64_t)row1 - 1; };
    auto row_of = [&](int64_t q) -> int32_t { return backward ? (int32_t)(n - q) : (int32_t)(q + 1); };
    // Chunk length R.  Row (v, p) runs in trip p + sigma * v, so an entry that reads (v2, p2) in an earlier
    // chunk needs sigma * (v - v2) > p2 - p; entries of the own chunk are earlier positions.  Any R gives a
    // valid schedule with sigma large enough; the best one has few trips, R + sigma * (C - 1).  On a grid
    // in natural ordering that is R = the grid's row length (five-point stencil: sigma = 1, nine-point:
    // sigma = 2), which is one of the offsets |i - j| the pattern uses; the bandwidth is always tried.
    int64_t bw = 1;
    int32_t S_all = 0;
    std::vector<int64_t> offs;           // distinct offsets, while there are few of them
    bool few = true;
    for (int32_t i = 1; i <= n; i++) {
        S_all = std::max(S_all, ptr1[i] - ptr1[i - 1]);
        for (int32_t k = ptr1[i - 1] - 1; k < ptr1[i] - 1; k++) {
            const int64_t d = pos_of(i) - pos_of(node1[k]);
            if (d <= 0) return;                      // not strictly triangular in sweep order: no plan
            bw = std::max(bw, d);
            if (few && std::find(offs.begin(), offs.end(), d) == offs.end()) {
                offs.push_back(d);
                if (offs.size() > 12) few = false;
            }
        }
    }
    const int64_t r_min = ((int64_t)n + kSweepMaxChunks - 1) / kSweepMaxChunks;
    std::vector<int64_t> cand{std::max(bw, r_min)};
    if (few)
        for (int64_t d : offs)
            if (d > 1 && d >= r_min && std::find(cand.begin(), cand.end(), d) == cand.end()) cand.push_back(d);
    auto sigma_for = [&](int64_t R_) {
        int64_t sg = 1;
        for (int32_t i = 1; i <= n; i++) {
            const int64_t q = pos_of(i), v = q / R_, p = q % R_;
            for (int32_t k = ptr1[i - 1] - 1; k < ptr1[i] - 1; k++) {
                const int64_t q2 = pos_of(node1[k]), v2 = q2 / R_, p2 = q2 % R_;
                if (v2 < v && p2 >= p) sg = std::max(sg, (p2 - p) / (v - v2) + 1);
            }
        }
        return sg;
    };
    int64_t R = 0, sigma = 0, best = -1;
    for (int64_t R_ : cand) {
        const int64_t sg = sigma_for(R_), C_ = ((int64_t)n + R_ - 1) / R_, tr = R_ + sg * (C_ - 1);
        if (best < 0 || tr < best) { best = tr; R = R_; sigma = sg; }
    }
    const int64_t C = ((int64_t)n + R - 1) / R;
    const int64_t trips = R + sigma * (C - 1);
    P.R = (int32_t)R;
    P.sigma = (int32_t)sigma;
    P.C = (int32_t)C;
    P.trips = (int32_t)std::min<int64_t>(trips, INT32_MAX);
    P.S_max = S_all;
    // worth it?  a trip costs a fraction of a microsecond, a level launch ~4 us; a shallow schedule is
    // better served by the level launches, a sweep with fewer than a warp of chunks has no wavefront
    if (S_all > 254 || C < 32 || levels <= 64 || trips > 16 * levels || trips > (int64_t)1 << 24) return;

    // trips: active chunk range, sizes, offsets
    P.trip.resize((size_t)trips);
    int64_t off = 0, soff = 0;
    int32_t w16_max = 0;
    std::vector<int32_t> S_of((size_t)trips, 0);
    for (int32_t i = 1; i <= n; i++) {
        const int64_t q = pos_of(i), v = q / R, p = q % R;
        int32_t &s = S_of[(size_t)(p + sigma * v)];
        s = std::max(s, ptr1[i] - ptr1[i - 1]);
    }
    for (int64_t t = 0; t < trips; t++) {
        const int64_t vlo = std::max<int64_t>(0, (t - (R - 1) + sigma - 1) / sigma);    // ceil((t - R + 1) / sigma)
        const int64_t vhi = std::min<int64_t>(C - 1, t / sigma);
        SweepTrip &T = P.trip[(size_t)t];
        T.vlo = (int32_t)vlo;
        T.w = (int32_t)std::max<int64_t>(0, vhi - vlo + 1);
        T.w16 = (T.w + 15) & ~15;
        T.S = S_of[(size_t)t];
        T.off = off;
        T.soff = soff;
        off += T.w16;
        soff += (int64_t)T.S * T.w16;
        w16_max = std::max(w16_max, T.w16);
    }
    P.total = off;
    P.total_s = soff;
    P.w16_max = w16_max;
    if (soff > 8 * ne + (int64_t)65536 || off > (int64_t)INT32_MAX) return;   // padding blow-up / far index range

    // ring depth: how many positions back the readers of a chunk look (distance in the OWNER's positions)
    int64_t dmax = 1;
    for (int32_t i = 1; i <= n; i++) {
        const int64_t q = pos_of(i), v = q / R, p = q % R, t = p + sigma * v;
        for (int32_t k = ptr1[i - 1] - 1; k < ptr1[i] - 1; k++) {
            const int64_t q2 = pos_of(node1[k]), v2 = q2 / R, p2 = q2 % R;
            const int64_t d = (t - sigma * v2) - p2;
            if (d < 1 || v2 > v) return;                           // cannot happen (sigma as above)
            dmax = std::max(dmax, d);
        }
    }
    // shared memory: nstage stages of [rhs | val | src | cnt] + the ring; deep rings give way to stages
    const int64_t stage = (((int64_t)w16_max * (8 + 12 * (int64_t)S_all + 1)) + 127) & ~(int64_t)127;
    // (depth a power of two: the kernel masks instead of dividing; entries further back than the ring are
    //  read from the trip-ordered solution in global memory)
    int64_t W = 2;
    while (W < dmax + 1 && W < 8) W *= 2;
    int64_t nstage = 0;
    for (; W >= 2; W /= 2) {
        const int64_t ring = W * C * 8;
        nstage = std::min<int64_t>(4, (kSweepSmemBudget - ring) / stage);
        if (nstage >= 2) break;
    }
    if (W < 2 || nstage < 2) return;
    P.has_far = dmax + 1 > W ? 1 : 0;
    P.W = (int32_t)W;
    P.nstage = (int32_t)nstage;
    P.stage_bytes = (int32_t)stage;
    P.threads = (int32_t)std::min<int64_t>(1024, (C + 31) & ~(int64_t)31);

    // slots: ldu.cu fills them on the device from the device-resident factor pattern (sweep_slots_kernel, the
    // same statements); the host copy is for the tests
    if (!fill_slots) {
        P.eligible = true;
        return;
    }
    P.src.assign((size_t)soff, 0);
    P.valmap.assign((size_t)soff, -1);
    P.cnt.assign((size_t)off, kSweepNoRow);
    for (int64_t t = 0; t < trips; t++) {
        const SweepTrip &T = P.trip[(size_t)t];
        for (int32_t u = 0; u < T.w; u++) {
            const int64_t v = T.vlo + u, p = t - sigma * v, q = v * R + p;
            if (q >= n) continue;                                  // the last chunk may be short
            const int32_t i = row_of(q);
            const int32_t kb = ptr1[i - 1] - 1, c = ptr1[i] - 1 - kb;
            P.cnt[(size_t)(T.off + u)] = (uint8_t)c;
            for (int32_t s = 0; s < c; s++) {
                const int64_t q2 = pos_of(node1[kb + s]), v2 = q2 / R, p2 = q2 % R, t2 = p2 + sigma * v2;
                const int64_t d = (t - sigma * v2) - p2;           // >= 1: finished in an earlier trip
                const size_t slot = (size_t)(T.soff + (int64_t)s * T.w16 + u);
                P.valmap[slot] = kb + s;
                if (d < W) {
                    P.src[slot] = (int32_t)((p2 % W) * C + v2);
                } else {
                    const SweepTrip &T2 = P.trip[(size_t)t2];
                    P.src[slot] = -(int32_t)(1 + T2.off + (v2 - T2.vlo));
                }
            }
        }
    }
    P.eligible = true;
}

}  // namespace sigb

extern "C" {

// Diagnostic / test entry (no GPU): the plan of a statically scheduled sweep.  info[16] = eligible, R,
// sigma, C, trips, W, S_max, w16_max, nstage, stage_bytes, threads, total (lo, hi), total_s (lo, hi), has_far.
// The arrays may be null (first call: sizes); trip_table is trips x 8 int32 (vlo, w, w16, S, off lo / hi,
// soff lo / hi), src and valmap total_s, cnt total.
int sigb_debug_ldu_sweep_plan(int32_t n, const int32_t *ptr1, const int32_t *node1, int backward, int64_t levels,
                              int32_t *info, int32_t *trip_table, int32_t *src, uint8_t *cnt, int64_t *valmap)
{
    SIGB_REQUIRE(n >= 0 && ptr1 && info && (n == 0 || ptr1[n] == 1 || node1), SIGB_ERR_ARG,
                 "sigb_debug_ldu_sweep_plan: bad argument");
    SweepPlan P;
    build_sweep_plan(n, ptr1, node1, backward, levels, P, true);
    const int32_t vals[16] = {P.eligible ? 1 : 0, P.R, P.sigma, P.C, P.trips, P.W, P.S_max, P.w16_max, P.nstage,
                              P.stage_bytes, P.threads, (int32_t)(P.total & 0xffffffff), (int32_t)(P.total >> 32),
                              (int32_t)(P.total_s & 0xffffffff), (int32_t)(P.total_s >> 32), P.has_far};
    for (int k = 0; k < 16; k++) info[k] = vals[k];
    if (!P.eligible) return SIGB_OK;
    if (trip_table)
        for (size_t t = 0; t < P.trip.size(); t++) {
            const SweepTrip &T = P.trip[t];
            int32_t *o = trip_table + 8 * t;
            o[0] = T.vlo; o[1] = T.w; o[2] = T.w16; o[3] = T.S;
            o[4] = (int32_t)(T.off & 0xffffffff); o[5] = (int32_t)(T.off >> 32);
            o[6] = (int32_t)(T.soff & 0xffffffff); o[7] = (int32_t)(T.soff >> 32);
        }
    if (src) std::copy(P.src.begin(), P.src.end(), src);
    if (cnt) std::copy(P.cnt.begin(), P.cnt.end(), cnt);
    if (valmap) std::copy(P.valmap.begin(), P.valmap.end(), valmap);
    return SIGB_OK;
}

}  // extern "C"
