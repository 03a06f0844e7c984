// Sliding-window bilinear demosaic on packed bytes (debayer.cpp:45-79 == cv::demosaicing + R/B swap): a thread owns four
// adjacent columns and walks down the rows, keeping per Bayer row what its neighbours need from it.
#pragma once
#include "frame_math.cuh"

namespace rip {

// ---- vertical sliding window over one 4-pixel column of the staged tile ------------------------------
// A thread that walks consecutive rows reuses everything a row contributes to its neighbours: per Bayer row the
// three packed words (centre / shifted left / shifted right) and, in 16-bit lanes, what the rows above and below
// need from it (A, S) and what it needs from itself (W).  Per output row that leaves 3 shared-memory loads, 2 funnel
// shifts and 5 byte permutes instead of 9 / 6 / 8 (frame_math.cuh demosaic_quad_swar is the reference form).
struct BayerRow {
  uint32_t c, l, r;  // columns x..x+3, x-1..x+2, x+1..x+4
  uint32_t A;        // centre word, lanes at the colour sites of the rows above / below
  uint32_t S;        // left + right words, same lanes (their diagonal contribution)
  uint32_t W;        // left + right words, lanes at this row's own colour sites, + rounding constant
};
// `img_row`: row index in the frame (decides the CFA phase); `srow`: row index in the staged tile
// `p`: this lane's position in the staged row = the word holding columns x-4 .. x-1
__device__ __forceinline__ BayerRow load_bayer_row(const uint32_t* p, int img_row, int cfa) {
  const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
  BayerRow b;
  b.c = w1; b.l = funnel_r(w0, w1, 24); b.r = funnel_r(w1, w2, 8);
  const uint32_t cpar = (uint32_t)((cfa ^ (cfa >> 1) ^ img_row) & 1);  // column parity of this row's colour sites
  const uint32_t own = 0x4240u + 0x0101u * cpar, other = 0x4341u - 0x0101u * cpar;
  b.A = prmt(b.c, 0u, other);
  b.S = prmt(b.l, 0u, other) + prmt(b.r, 0u, other);
  b.W = prmt(b.l, 0u, own) + prmt(b.r, 0u, own) + 0x00020002u;
  return b;
}
// packed B / G / R of the row `m` (frame row img_row) between rows `n` (above) and `s` (below)
__device__ __forceinline__ void demosaic_window(const BayerRow& n, const BayerRow& m, const BayerRow& s, int img_row, int cfa,
                                                uint32_t& Bw, uint32_t& Gw, uint32_t& Rw) {
  const int cpar = (cfa ^ (cfa >> 1) ^ img_row) & 1;
  const bool row_has_r = (((img_row ^ (cfa >> 1)) & 1) == 0);
  const uint32_t H = avg_round_u8x4(m.l, m.r), V = avg_round_u8x4(n.c, s.c);
  const uint32_t X = ((n.A + s.A + m.W) >> 2) & 0x00ff00ffu;
  const uint32_t D = ((n.S + s.S + 0x00020002u) >> 2) & 0x00ff00ffu;
  const uint32_t sel = 0x7250u - 0x4c4cu * (uint32_t)cpar;
  const uint32_t site = 0x00ff00ffu << (8 * cpar);
  Gw = prmt(X, m.c, sel);
  const uint32_t row_colour = (m.c & site) | (H & ~site);
  const uint32_t other_colour = prmt(D, V, sel);
  Rw = row_has_r ? row_colour : other_colour;
  Bw = row_has_r ? other_colour : row_colour;
}

// ---- the same with everything that depends on a row's CFA phase precomputed ---------------------------------------
// Rows alternate between two phases.  A kernel that walks rows computes one BayerPhase for its "even" rows (the rows
// with the parity of a reference row) once; what odd rows need follows from it at no cost: their lane selectors are the
// even ones swapped, their site mask the complement, their R/B assignment the opposite -- only the merge selector of odd
// rows takes a register of its own.  With the row loop unrolled by an even count, no per-row selector arithmetic is left.
struct BayerPhase {
  uint32_t own, other;    // 16-bit lane selectors of an even row's own colour sites / of its other columns
  uint32_t sel_e, sel_o;  // merge selectors with an even / odd row as the centre
  uint32_t site;          // byte mask of an even row's colour sites
  bool row_has_r;         // an even row holds R (and G) sites
};
__device__ __forceinline__ BayerPhase bayer_phase(int even_row, int cfa) {
  const uint32_t cpar = (uint32_t)((cfa ^ (cfa >> 1) ^ even_row) & 1);
  BayerPhase q;
  q.own = 0x4240u + 0x0101u * cpar; q.other = 0x4341u - 0x0101u * cpar;
  q.sel_e = 0x7250u - 0x4c4cu * cpar; q.sel_o = 0x2604u + 0x4c4cu * cpar;
  q.site = 0x00ff00ffu << (8 * cpar);
  q.row_has_r = (((even_row ^ (cfa >> 1)) & 1) == 0);
  return q;
}
template <bool ODD>
__device__ __forceinline__ BayerRow load_bayer_row(const uint32_t* p, const BayerPhase& q) {
  const uint32_t own = ODD ? q.other : q.own, other = ODD ? q.own : q.other;
  const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
  BayerRow b;
  b.c = w1; b.l = funnel_r(w0, w1, 24); b.r = funnel_r(w1, w2, 8);
  b.A = prmt(b.c, 0u, other);
  b.S = prmt(b.l, 0u, other) + prmt(b.r, 0u, other);
  b.W = prmt(b.l, 0u, own) + prmt(b.r, 0u, own) + 0x00020002u;
  return b;
}
// ODD: the centre row `m` is an odd row
template <bool ODD>
__device__ __forceinline__ void demosaic_window(const BayerRow& n, const BayerRow& m, const BayerRow& s, const BayerPhase& q,
                                                uint32_t& Bw, uint32_t& Gw, uint32_t& Rw) {
  const uint32_t H = avg_round_u8x4(m.l, m.r), V = avg_round_u8x4(n.c, s.c);
  const uint32_t X = ((n.A + s.A + m.W) >> 2) & 0x00ff00ffu;
  const uint32_t D = ((n.S + s.S + 0x00020002u) >> 2) & 0x00ff00ffu;
  const uint32_t sel = ODD ? q.sel_o : q.sel_e;
  Gw = prmt(X, m.c, sel);
  const uint32_t row_colour = ODD ? ((m.c & ~q.site) | (H & q.site)) : ((m.c & q.site) | (H & ~q.site));
  const uint32_t other_colour = prmt(D, V, sel);
  const bool row_has_r = ODD ? !q.row_has_r : q.row_has_r;
  Rw = row_has_r ? row_colour : other_colour;
  Bw = row_has_r ? other_colour : row_colour;
}

// ---- the same for a rolled row loop: the phase of the current centre row is kept in registers and flipped per row ------
// (five one-cycle operations per row; used where unrolling the row loop would overflow the instruction cache)
struct BayerPhaseRT {
  uint32_t own, other;  // lane selectors of the CENTRE row (its bottom row uses them swapped)
  uint32_t sel, site;
  bool row_has_r;
  __device__ __forceinline__ void flip() {
    own ^= 0x0101u; other ^= 0x0101u;  // 0x4240 <-> 0x4341
    sel ^= 0x5454u;                    // 0x7250 <-> 0x2604
    site = ~site;
    row_has_r = !row_has_r;
  }
};
__device__ __forceinline__ BayerPhaseRT bayer_phase_rt(int centre_row, int cfa) {
  const BayerPhase q = bayer_phase(centre_row, cfa);
  BayerPhaseRT r;
  r.own = q.own; r.other = q.other; r.sel = q.sel_e; r.site = q.site; r.row_has_r = q.row_has_r;
  return r;
}
// the row BELOW a centre row of phase q
__device__ __forceinline__ BayerRow load_bayer_row_below(const uint32_t* p, const BayerPhaseRT& q) {
  const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
  BayerRow b;
  b.c = w1; b.l = funnel_r(w0, w1, 24); b.r = funnel_r(w1, w2, 8);
  b.A = prmt(b.c, 0u, q.own);
  b.S = prmt(b.l, 0u, q.own) + prmt(b.r, 0u, q.own);
  b.W = prmt(b.l, 0u, q.other) + prmt(b.r, 0u, q.other) + 0x00020002u;
  return b;
}
__device__ __forceinline__ void demosaic_window(const BayerRow& n, const BayerRow& m, const BayerRow& s, const BayerPhaseRT& q,
                                                uint32_t& Bw, uint32_t& Gw, uint32_t& Rw) {
  const uint32_t H = avg_round_u8x4(m.l, m.r), V = avg_round_u8x4(n.c, s.c);
  const uint32_t X = ((n.A + s.A + m.W) >> 2) & 0x00ff00ffu;
  const uint32_t D = ((n.S + s.S + 0x00020002u) >> 2) & 0x00ff00ffu;
  Gw = prmt(X, m.c, q.sel);
  const uint32_t row_colour = (m.c & q.site) | (H & ~q.site);
  const uint32_t other_colour = prmt(D, V, q.sel);
  Rw = q.row_has_r ? row_colour : other_colour;
  Bw = q.row_has_r ? other_colour : row_colour;
}

}  // namespace rip
