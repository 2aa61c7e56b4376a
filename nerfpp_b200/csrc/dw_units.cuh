// Work table of the weight-gradient kernel (mlp_nerf_bwd_dw_kernel, mlp_nerf_bwd_tc.cu), shared by the classic-NeRF backward and the LeRF
// head backward (lerf_bwd_tc.cu).  A unit is one product D[M = columns of the A operand (128 or 256), N = columns of the B operand] = A^T B
// summed over the rows of all tiles; both operands are column ranges of per-tile record regions ([row half][column chunk][64 rows][8] bf16,
// mlp_nerf_layout.cuh / lerf_layout.cuh) as they lie in HBM, fetched with one bulk copy per operand and 64-row slab.
#pragma once
#include "common.cuh"

namespace nrf {
namespace dw {

constexpr int kMaxUnits = 16;
struct Unit {
	int32_t a_src, a_off, a_cols;       // record kind (0 gradient, 1 saved), byte offset of the operand's first column chunk in the record, operand width (128 / 256)
	int32_t b_src, b_off, b_cols;       // b_cols: operand width = MMA N (multiple of 16, <= 256)
	int32_t n_lo, n_hi;                 // D(m, n) is written for n_lo <= n < n_hi ...
	int32_t stride_m, stride_n;         // ... to out[m * stride_m + (n - n_lo) * stride_n]
	int32_t bias_mode;                  // 0 none; 1 column sums of A -> bias[m]; 2 column sums of B -> bias[0..2] (rgb), bias2[0] (alpha)
	int32_t a_half, b_half;             // bytes between the two 64-row halves of the operand's region (0: a_cols / b_cols x 128, i.e. the operand IS the region)
	int32_t n2_lo, n2_hi;               // optional second output range: n2_lo <= n < n2_hi goes to out2[m * stride_m + (n - n2_lo) * stride_n]
	int32_t pad_;
	float* out;
	float* bias;
	float* bias2;
	float* out2;
};
struct UnitTable {
	Unit u[kMaxUnits];
	int32_t count;
	int32_t pad_;
	int64_t save_tile, grad_tile;       // bytes per 128-row tile of the two record kinds
};

// one persistent CTA per SM over the flat (unit, 64-row slab) list; accumulators leave as fp32 REDs into the units' outputs (+=)
int launch_dw_units(const UnitTable& T, const void* saved, const void* grads, int64_t n_tiles, nrf_stream stream);

}  // namespace dw
}  // namespace nrf
