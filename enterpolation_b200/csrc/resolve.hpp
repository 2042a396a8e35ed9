// Host-side "builder" of the engine: validates an enterp_curve_desc with the reference's builder
// rules and materialises the tables the kernels read. Product code (no oracle involvement).
//
// Unlike the reference, which keeps knot chains lazy (Equidistant recomputes step*i+offset at every
// access, BorderBuffer/BorderDeletion remap indices at every access), the engine TABULATES the
// virtual knot vector once at build time. Each table entry is produced by the very same IEEE
// operations the reference would execute per access (this file is compiled with
// -ffp-contract=off), so the tabulated values are bit-identical to the recomputed ones.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/enterp.h"

namespace enterp {

enum Adaptor : int32_t { AD_NONE = 0, AD_BORDER_BUFFER = 1, AD_BORDER_DELETION = 2 };
enum Stage : int32_t { STAGE_ELEMENTS = 0, STAGE_KNOTS = 1, STAGE_SPACE = 2 };

struct SearchLevels {
  // Pivot tree over the searched inner knot range, top level first; nodes are 32-byte sectors
  // (8 f32 / 4 f64 keys) and every level is NaN padded (NaN <= x is false, so padding never counts).
  std::vector<std::vector<unsigned char>> level;  // raw R values
};

struct Resolved {
  int32_t kind = 0, scalar = 0, dim = 0, weighted = 0;
  int32_t dh = 0;          // components carried through the recursion: dim + weighted
  int32_t knot_base = -1;  // ENTERP_KNOTS_SORTED | ENTERP_KNOTS_EQUIDISTANT | -1 (Bezier)
  int32_t adaptor = AD_NONE;
  uint64_t n_elem = 0;
  uint64_t border_n = 0;
  uint64_t inner_len = 0;  // innermost knot chain length
  uint64_t virt_len = 0;   // K::len()
  uint64_t degree = 0;
  uint64_t space_len = 0;
  // knot store: SORTED+none/EQUI+none: virtual == inner == store; BorderBuffer: store = virtual
  // (inner j at store[j + border_n]); BorderDeletion: store = inner (virtual i at store[i + 1]).
  std::vector<unsigned char> store;  // raw R
  uint64_t n_store = 0;
  uint32_t vk_off = 0, ik_off = 0;
  // span search bounds on the inner chain and the map back to virtual indices
  uint64_t imin = 0, imax = 0;
  int64_t map_add = 0;
  uint64_t map0 = 0, mapL = 0;
  double eq_step = 0, eq_offset = 0, eq_first = 0;  // exact R values widened to double
  // elements, lifted to homogeneous form when weighted: [n_elem][dh]
  std::vector<unsigned char> elems;
  SearchLevels levels;
  // Bucket index over the searched knot range (sorted knots): n_buckets equal-width buckets of [first, last], stored
  // as one 16-byte word per block of 32 buckets: { knots in all lower blocks, 32 x 2-bit knot counts (two words),
  // flag: a bucket of the block holds more than 3 knots }. bucket(x) is the monotone function
  // clamp(trunc((x - first) * inv), 0, n_buckets - 1) evaluated in R, so the knots <= x are those of the lower buckets
  // plus a compare over the few knots sharing x's bucket. Empty when the knots are too clustered for it.
  std::vector<uint32_t> buckets;
  uint32_t n_buckets = 0;
  double bkt_first = 0, bkt_inv = 0;
  double domain[2] = {0, 0};
  // Bezier TransformInput (base/adaptors.rs:137-139)
  int32_t has_transform = 0;
  double in_add = 0, in_mul = 1;
  // Generic B-spline kernel, degree 1..7: the 2*degree knots of every span index in [lo, hi], padded to
  // 32 bytes, so that they arrive with one or two 256-bit loads instead of 2*degree scalar ones
  std::vector<unsigned char> knot_rows;
  uint32_t krow_stride = 0;  // in units of R; 0 = no table
  // Linear over sorted knots: one record per knot interval j = [k_j, k_{j+1}, e_j[0..dh), e_{j+1}[0..dh)],
  // padded to a multiple of 16 bytes, so that the lerp's operands arrive in one or two 32-byte sectors
  std::vector<unsigned char> records;
  uint32_t rec_stride = 0;  // in units of R; 0 = no record table
  // Linear over sorted knots, batches: "slot" table = the interval records again, but indexed by an equal-width grid
  // over [first, last] instead of by interval: slot b holds the record of the interval that contains the START of
  // grid cell b, i.e. interval P_b - 1 with P_b = number of knots whose cell is below b (same monotone cell function
  // as the bucket index, own resolution), clamped to [0, len - 2]; n_slots + 1 entries. One gather then serves every
  // sample with no knot between its cell's start and itself, a second one (slot b + 1) those beyond the cell's last
  // knot; the rest (two or more knots in the cell and x between them) falls back to index + records.
  std::vector<unsigned char> slots;
  uint32_t n_slots = 0;
  double slot_inv = 0;
  // B-splines over >= 4096 sorted knots, generic kernel: the knot rows again, indexed by the same kind of grid (2 cells
  // per searched knot): slot b = knot row of span imin + P_b (+ map_add), its last (padding) value carrying that span
  // index as raw bits. Only where a row has padding to spare (2 * degree < krow_stride).
  std::vector<unsigned char> kslots;
  uint32_t n_kslots = 0;
  double kslot_inv = 0;
  // Linear easing (easing/mod.rs:86-132, plateau.rs): kind, N of smoothstart/smoothend, Plateau min/max in R
  int32_t easing = 0, easing_n = 0;
  double ease_min = 0, ease_max = 1;
  // span-row table of the fast B-spline kernel (rows.cuh / rows_layout.hpp); rows_mode 0 = none,
  // 1 = every reachable denominator is a power of two (exact inverse multiply), 2 = exact division
  // from tabulated reciprocals.
  int32_t rows_mode = 0;
  uint32_t rows_n = 0, rows_lo = 0, rows_len = 0;
  uint32_t rows_pow2_ops = 0;  // exact-division rows: factor operations whose denominators are powers of two in every row
  std::vector<unsigned char> rows;
  // equidistant span division (x - offset) / step: 0 IEEE divide, 1 exact inverse multiply, 2 exact
  // division from eq_rstep = RN(1 / step)
  int32_t eq_mode = 0;
  double eq_rstep = 0;
};

// Validates up to and including `stage`; fills `err` like the reference's error structs.
enterp_status validate(const enterp_curve_desc& d, int stage, enterp_error_info* err);
// Full validation + table construction.
enterp_status resolve(const enterp_curve_desc& d, Resolved& out, enterp_error_info* err);

}  // namespace enterp
