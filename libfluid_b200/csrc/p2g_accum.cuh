// Used by the marching P2G kernel (p2g_march.cu): the per-cell register accumulation of one velocity
// component from particles staged in shared memory as [field][cell * PB_CSTRIDE + slot].
// Reference: simulation::_transfer_to_grid_{pic,flip,apic}, src/simulation.cpp:293-412.
#pragma once
#include "lfk_internal.cuh"

#define PB_WIN 4                 // particle slots per cell staged at a time
#define PB_CSTRIDE (PB_WIN + 1)  // padded: lane stride of 5 doubles is bank-conflict free
#define PB_FSTRIDE (32 * PB_CSTRIDE)
#define PB_FIELDS 7              // pos(3) + v_k + c_k(3)

struct PBParams {
	double half, inv_h;
	double gdt[3];
	int add_gravity;
};

__device__ __forceinline__ void cp_async8(double *smem, const double *gmem) {
	unsigned s = (unsigned)__cvta_generic_to_shared(smem);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
	asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
// hat weight max(0, 1 - |d|) (reference _kernel, src/simulation.cpp:207-213).  fmax(double, double) costs ~7
// instructions on sm_100 (DSETP on the fp64 pipe, selects on both halves, NaN fix-up) and the eight weights per particle
// and component were half of the kernel's instructions (profiles/r1c_p2g_march_hotloop.txt); here the clamp is a sign
// test on the high word and two selects on the integer pipe, which leaves the fp64 pipe (the kernel's limiter: one warp
// instruction per two cycles) to the sums.  Exact: the result is u or +0.
// SCALE (PIC / FLIP, src/simulation.cpp:313-315): the weight argument is d / h; APIC uses d itself (:367-369).
template <bool SCALE> __device__ __forceinline__ double hatw(double d, double inv_h) {
	if (SCALE) { d *= inv_h; } // h == 1: inv_h == 1 and the product is d itself
	const double u = 1.0 - fabs(d);
	const int hi = __double2hiint(u), lo = __double2loint(u);
	return __hiloint2double(hi < 0 ? 0 : hi, hi < 0 ? 0 : lo);
}

// Contributions of the staged particles of one cell to one velocity component.  COMP selects which axis is
// staggered: the staggered axis has 2 reachable faces (cell - 1, cell), the other two axes 3 (cell - 1 .. cell + 1).
template <int COMP, bool APIC> __device__ __forceinline__ void accumulate_cell(const double *__restrict__ st,
	int lane, int nslots, const double *cc /* cell-centre coords of cell-1, cell, cell+1 per axis: [3][3] */,
	double half, double inv_h, double *accw, double *accv) {
	constexpr bool SCALE = !APIC;
	constexpr int NA = COMP == 0 ? 2 : 3, NB = COMP == 1 ? 2 : 3, NC = COMP == 2 ? 2 : 3;
	// sample positions per axis: staggered axis -> +face of (cell - 1), +face of cell; others -> centres
	double sx[NA], sy[NB], sz[NC];
#pragma unroll
	for (int a = 0; a < NA; ++a) { sx[a] = COMP == 0 ? cc[0 * 3 + a] + half : cc[0 * 3 + a]; }
#pragma unroll
	for (int b = 0; b < NB; ++b) { sy[b] = COMP == 1 ? cc[1 * 3 + b] + half : cc[1 * 3 + b]; }
#pragma unroll
	for (int c = 0; c < NC; ++c) { sz[c] = COMP == 2 ? cc[2 * 3 + c] + half : cc[2 * 3 + c]; }
	const double *sp = st + lane * PB_CSTRIDE;
	for (int s = 0; s < nslots; ++s) {
		const double px = sp[0 * PB_FSTRIDE + s], py = sp[1 * PB_FSTRIDE + s], pz = sp[2 * PB_FSTRIDE + s];
		const double vk = sp[3 * PB_FSTRIDE + s];
		double c0 = 0.0, c1 = 0.0, c2 = 0.0;
		if (APIC) {
			c0 = sp[4 * PB_FSTRIDE + s];
			c1 = sp[5 * PB_FSTRIDE + s];
			c2 = sp[6 * PB_FSTRIDE + s];
		}
		double wx[NA], wy[NB], wz[NC], ax[NA], by[NB], cz[NC];
#pragma unroll
		for (int a = 0; a < NA; ++a) {
			double d = px - sx[a];
			wx[a] = hatw<SCALE>(d, inv_h);
			ax[a] = APIC ? vk - c0 * d : vk; // v_k + c_k0 * (x_sample - x_p)
		}
#pragma unroll
		for (int b = 0; b < NB; ++b) {
			double d = py - sy[b];
			wy[b] = hatw<SCALE>(d, inv_h);
			by[b] = APIC ? -c1 * d : 0.0;
		}
#pragma unroll
		for (int c = 0; c < NC; ++c) {
			double d = pz - sz[c];
			wz[c] = hatw<SCALE>(d, inv_h);
			cz[c] = APIC ? -c2 * d : 0.0;
		}
		// sum(w) and sum(w (ax + by + cz)) with w = wx wy wz, as three fused multiply-adds per face:
		//   w (ax + by + cz) = (wy wz) (wx ax) + wx ((wy wz) (by + cz))
		double wax[NA];
#pragma unroll
		for (int a = 0; a < NA; ++a) { wax[a] = wx[a] * ax[a]; }
#pragma unroll
		for (int c = 0; c < NC; ++c) {
#pragma unroll
			for (int b = 0; b < NB; ++b) {
				const double wyz = wy[b] * wz[c], wbc = APIC ? wyz * (by[b] + cz[c]) : 0.0;
#pragma unroll
				for (int a = 0; a < NA; ++a) {
					const int t = (c * NB + b) * NA + a;
					accw[t] = fma(wx[a], wyz, accw[t]);
					accv[t] = APIC ? fma(wax[a], wyz, fma(wx[a], wbc, accv[t])) : fma(wax[a], wyz, accv[t]);
				}
			}
		}
	}
}

