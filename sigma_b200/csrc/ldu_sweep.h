// ldu_sweep.h -- the statically scheduled ("systolic") triangular sweeps of the ILDU(0)
// preconditioner: plan built on the host (ldu_host.cpp, no CUDA), run by ldu.cu.
//
// A sweep (I + M) x = rhs with M strictly triangular in sweep order is cut into C chunks of R
// consecutive positions.  Row (v, p) -- chunk v, position p -- is computed in trip p + sigma * v: all
// chunks advance together, each sigma trips behind its predecessor, with sigma just large enough for
// every stored dependency to be finished one trip earlier at the latest (R = N, sigma = 1 for the
// five-point stencil on an N x N grid in natural ordering: the classic wavefront; sigma = 2 for the
// nine-point stencil).  R is chosen among the offsets the pattern uses and its bandwidth for the fewest
// trips, R + sigma (C - 1).  The whole sweep is ONE CTA: one thread per
// chunk, one barrier per trip, finished values handed on through a shared-memory ring of the last
// few positions of every chunk; everything a trip reads is stored as one contiguous slab per trip,
// which the TMA engine brings into shared memory a few trips ahead.  Nothing
// is polled: the schedule is proven valid on the host for the pattern at hand, otherwise the plan
// is not eligible and ldu.cu keeps its other forms.
#pragma once

#include <cstdint>
#include <vector>

namespace sigb {

struct SweepTrip {       // one trip: chunks [vlo, vlo + w) are active
    int32_t vlo, w;
    int32_t w16;         // w rounded up to 16 (the stride of the slot-major arrays of this trip)
    int32_t S;           // slots = longest row of the trip
    int64_t off;         // start of the trip in rhs / x / cnt (elements; multiple of 16)
    int64_t soff;        // start of the trip in val / src (elements; slot s of chunk vlo + u at soff + s * w16 + u)
};

struct SweepPlan {
    bool eligible = false;
    int32_t n = 0, backward = 0;
    int32_t R = 0, sigma = 0, C = 0, trips = 0;
    int32_t W = 0;                 // ring depth (positions kept per chunk)
    int32_t S_max = 0, w16_max = 0;
    int32_t nstage = 0, stage_bytes = 0, threads = 0;
    int32_t has_far = 0;           // some entries are read from the trip-ordered solution in global memory
    int64_t total = 0, total_s = 0;
    std::vector<SweepTrip> trip;
    std::vector<int32_t> src;      // >= 0: ring index (p' mod W) * C + v' ; < 0: -(1 + index into the trip-ordered x)
    std::vector<uint8_t> cnt;      // entries of the row; 0xFF: no row at this (trip, chunk)
    std::vector<int64_t> valmap;   // entry of the factor's value array behind every slot, -1 = padding
                                   // (src / cnt / valmap: filled on request only -- the device builds its own)
};

constexpr int kSweepMaxChunks = 4096;
constexpr int kSweepSmemBudget = 216 * 1024;   // stages + ring (the mbarriers and the runtime's 1 KB come on top)
constexpr uint8_t kSweepNoRow = 0xFF;

// ptr1 / node1: the rows of the strictly triangular factor (1-based).  backward = 0: positions are rows
// ascending (L); 1: rows descending (U).  levels: depth of the level schedule of the same sweep.
// fill_slots = false: schedule and trip table only (src / cnt / valmap stay empty).
void build_sweep_plan(int32_t n, const int32_t *ptr1, const int32_t *node1, int backward, int64_t levels,
                      SweepPlan &P, bool fill_slots);

}  // namespace sigb
