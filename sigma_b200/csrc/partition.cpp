// partition.cpp -- host-side index work of the row-sharded operators: row
// partition balanced by stored entries, halo lists, local renumbering.
// Pure int32 work with no CUDA call, so it runs (and is tested) without a GPU;
// its results must equal oracle/sigma_oracle.c (orc_partition_rows,
// orc_halo_build) bit for bit.
//
// The reference is serial; the seam these lists generalise is the block-row
// loop of composite_matvec_add, src/matrix/sparse_matrix_composites.f90:1076-1100
// (x(j1:j2) / y(i1:i2) slices per block).
#include <algorithm>
#include <vector>

#include "internal.h"

using namespace sigb;

extern "C" {

// part[r] = first row whose preceding entries reach r/nparts of all entries.
int sigb_partition_rows(int32_t n, const int32_t *ptr1, int32_t nparts, int32_t *part)
{
    SIGB_REQUIRE(ptr1 && part && n >= 0 && nparts >= 1, SIGB_ERR_ARG, "sigb_partition_rows: bad argument");
    const int64_t nnz = (int64_t)ptr1[n] - 1;
    int32_t i = 0;
    part[0] = 0;
    for (int32_t r = 1; r < nparts; r++) {
        const int64_t target = (nnz * r) / nparts;
        while (i < n && (int64_t)ptr1[i] - 1 < target) i++;
        part[r] = i;
    }
    part[nparts] = n;
    return SIGB_OK;
}

// Row tiling of the streaming CSR kernel (kernels_spmv.cu build_tiles_host), exposed so
// that the index work can be checked without a GPU.  tiles: 4 int32 per tile
// {first row, end row, first entry, end entry}, 0-based, capacity n tiles.
int sigb_debug_row_tiles(int32_t n, const int32_t *ptr1, int32_t *tiles, int32_t *ntiles)
{
    SIGB_REQUIRE(n >= 0 && ptr1 && tiles && ntiles, SIGB_ERR_ARG, "sigb_debug_row_tiles: bad argument");
    std::vector<TileDesc> t;
    build_tiles_host(ptr1, n, t);
    for (size_t k = 0; k < t.size(); k++) {
        tiles[4 * k + 0] = t[k].rs;
        tiles[4 * k + 1] = t[k].re;
        tiles[4 * k + 2] = t[k].ks;
        tiles[4 * k + 3] = t[k].ke;
    }
    *ntiles = (int32_t)t.size();
    return SIGB_OK;
}

// The balanced tiling of row-sharded operators (kernels_spmv.cu build_tiles_balanced), same layout.
int sigb_debug_row_tiles_balanced(int32_t n, const int32_t *ptr1, int32_t groups, int32_t *tiles, int32_t *ntiles)
{
    SIGB_REQUIRE(n >= 0 && ptr1 && tiles && ntiles, SIGB_ERR_ARG, "sigb_debug_row_tiles_balanced: bad argument");
    std::vector<TileDesc> t;
    build_tiles_balanced(ptr1, n, groups, t);
    for (size_t k = 0; k < t.size(); k++) {
        tiles[4 * k + 0] = t[k].rs;
        tiles[4 * k + 1] = t[k].re;
        tiles[4 * k + 2] = t[k].ks;
        tiles[4 * k + 3] = t[k].ke;
    }
    *ntiles = (int32_t)t.size();
    return SIGB_OK;
}

int sigb_halo_build(int32_t lo, int32_t hi, const int32_t *ptr_blk1, const int32_t *node_glob1,
                    int32_t *halo, int32_t *nhalo, int32_t *local_node)
{
    SIGB_REQUIRE(ptr_blk1 && nhalo && hi >= lo, SIGB_ERR_ARG, "sigb_halo_build: bad argument");
    const int32_t nloc = hi - lo;
    const int64_t cnt = (int64_t)ptr_blk1[nloc] - ptr_blk1[0];
    SIGB_REQUIRE(cnt == 0 || (node_glob1 && halo && local_node), SIGB_ERR_ARG, "sigb_halo_build: null array");
    std::vector<int32_t> h;
    h.reserve(1024);
    for (int64_t k = 0; k < cnt; k++) {
        const int32_t c = node_glob1[k];
        if (c <= lo || c > hi) h.push_back(c);
    }
    std::sort(h.begin(), h.end());
    h.erase(std::unique(h.begin(), h.end()), h.end());
    const int32_t nh = (int32_t)h.size();
    for (int32_t i = 0; i < nh; i++) halo[i] = h[i];
    for (int64_t k = 0; k < cnt; k++) {
        const int32_t c = node_glob1[k];
        if (c > lo && c <= hi) {
            local_node[k] = c - lo;
        } else {
            const int32_t pos = (int32_t)(std::lower_bound(h.begin(), h.end(), c) - h.begin());
            local_node[k] = nloc + 1 + pos;
        }
    }
    *nhalo = nh;
    return SIGB_OK;
}

}  // extern "C"
