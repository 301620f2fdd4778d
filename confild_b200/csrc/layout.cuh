// Packed-weight buffer layout shared by the pack kernels and every compute kernel.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/confild_cnf.h"

namespace cnf {

constexpr int kTileM = 128;           // query points per tensor-core tile (UMMA M)
constexpr int kSlabK = 64;            // 16-bit elements per 128-byte swizzled row
constexpr int kStageRows = 128;       // weight rows (UMMA N) per streamed stage
constexpr int kStageBytes = kStageRows * kSlabK * 2;  // 16 KiB

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline bool tc_shape_ok(int H) { return H == 128 || H == 256 || H == 384; }

// Stages per hidden layer: (H/128 row blocks) x (H/64 K slabs) x parts (hi, lo for the split).
__host__ __device__ inline int stages_per_layer(int H, int parts) { return (H / kStageRows) * (H / kSlabK) * parts; }

struct PackedLayout {
  // fp32 sections (offsets in bytes from the start of the packed buffer)
  size_t w_first;   // [H][cin]        w0 * net1.0.weight
  size_t w_out;     // [cout][H]       net1.{nl+1}.weight
  size_t b_out;     // [cout]          net1.{nl+1}.bias
  size_t b_shift;   // [(nl+1)*H]      w0 * net1.l.bias, l = 0..nl
  size_t v_cat;     // [(nl+1)*H][L]   w0 * net2.l.weight stacked over l
  size_t w_hid;     // [nl][H][H]      w0 * net1.l.weight (row n_out, col k_in), l = 1..nl
  size_t w_hid_t;   // [nl][H][H]      transposed copy (row k_in, col n_out)
  // tensor-core shared-memory images (only when tc_shape_ok)
  size_t tc_fwd_x3;  // [nl][stages_per_layer(H,2)][16 KiB] bf16 hi/lo of w0*W_l, operand B[n_out][k_in]
  size_t tc_fwd_h;   // [nl][stages_per_layer(H,1)][16 KiB] fp16
  size_t tc_bwd_x3;  // [nl][stages_per_layer(H,2)][16 KiB] bf16 hi/lo, operand B[k_in][n_out]
  // f16f8 precision: per (K slab, row block) an fp16 stage of S_l*w0*W_l and an fp8 stage whose 128-byte rows hold
  // [e4m3(S w)[k 0..63] | e4m3(S w - fp16(S w))[k 0..63]] of that slab; S_l = a power of two per layer
  size_t tc_fwd_f8;  // [nl][stages_per_layer(H,2)][16 KiB]
  size_t tc_scale;   // [2*nl] fp32: 1/S_l for l < nl, then S_l
  size_t w_first_t;  // [4][H] fp32: w_first coordinate-major, rows >= cin zero (16-byte loads in the kernels' layer 0)
  size_t total;
};

__host__ __device__ inline PackedLayout make_layout(const cnf_dims& d) {
  PackedLayout p{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  const size_t H = d.H, nl = d.nl;
  p.w_first = take(H * d.cin * 4);
  p.w_out = take((size_t)d.cout * H * 4);
  p.b_out = take((size_t)d.cout * 4);
  p.b_shift = take((nl + 1) * H * 4);
  p.v_cat = take((nl + 1) * H * d.L * 4);
  p.w_hid = take(nl * H * H * 4);
  p.w_hid_t = take(nl * H * H * 4);
  if (tc_shape_ok(d.H)) {
    p.tc_fwd_x3 = take(nl * (size_t)stages_per_layer(d.H, 2) * kStageBytes);
    p.tc_fwd_h = take(nl * (size_t)stages_per_layer(d.H, 1) * kStageBytes);
    p.tc_bwd_x3 = take(nl * (size_t)stages_per_layer(d.H, 2) * kStageBytes);
    p.tc_fwd_f8 = take(nl * (size_t)stages_per_layer(d.H, 2) * kStageBytes);
    p.tc_scale = take(2 * nl * sizeof(float));
  }
  p.w_first_t = take(4 * H * 4);
  p.total = off;
  return p;
}

// Offsets (in fp32 elements) of each tensor inside the flat parameter vector (state_dict order).
struct ParamOffsets {
  size_t w_first, b_first;  // net1.0
  size_t hid0;              // net1.1.weight; each hidden layer is H*H weights followed by H biases
  size_t w_out, b_out;      // net1.{nl+1}
  size_t v0;                // net2.0.weight; each is H*L
  size_t total;
};
__host__ __device__ inline ParamOffsets make_param_offsets(const cnf_dims& d) {
  ParamOffsets o{};
  const size_t H = d.H, nl = d.nl;
  o.w_first = 0;
  o.b_first = H * d.cin;
  o.hid0 = o.b_first + H;
  o.w_out = o.hid0 + nl * (H * H + H);
  o.b_out = o.w_out + (size_t)d.cout * H;
  o.v0 = o.b_out + d.cout;
  o.total = o.v0 + (nl + 1) * H * d.L;
  return o;
}

// Byte offset of element (row r in [0,128), k in [0,64)) inside one 16 KiB K-major SWIZZLE_128B block:
// 128-byte rows, 16-byte chunks XOR-ed with (row mod 8).  Used for weight stages (r = weight row) and
// for the activation operand (r = query point).
__host__ __device__ inline uint32_t sw128_offset(uint32_t r, uint32_t k) {
  return r * 128u + ((((k >> 3) ^ (r & 7u)) << 4) | ((k & 7u) << 1));
}

// Same layout for 8-bit elements: byte b in [0,128) of row r.
__host__ __device__ inline uint32_t sw128_byte_offset(uint32_t r, uint32_t b) {
  return r * 128u + ((((b >> 4) ^ (r & 7u)) << 4) | (b & 15u));
}

}  // namespace cnf
