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
