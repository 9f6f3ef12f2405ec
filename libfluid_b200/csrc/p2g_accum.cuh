// Shared by the P2G kernels (p2g_brick.cu, p2g_march.cu): the per-cell register accumulation of one velocity
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
	int hdiv; // PIC / FLIP: weights use (x_p - x_face) / h
};

__device__ __forceinline__ void cp_async8(double *smem, const double *gmem) {
	unsigned s = (unsigned)__cvta_generic_to_shared(smem);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
	asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ double hatw(double d) {
	return fmax(0.0, 1.0 - fabs(d));
}

// Contributions of the staged particles of one cell to one velocity component.  COMP selects which axis is
// staggered: the staggered axis has 2 reachable faces (cell - 1, cell), the other two axes 3 (cell - 1 .. cell + 1).
template <int COMP, bool APIC> __device__ __forceinline__ void accumulate_cell(const double *__restrict__ st,
	int lane, int nslots, const double *cc /* cell-centre coords of cell-1, cell, cell+1 per axis: [3][3] */,
	double half, double inv_h, int hdiv, double *accw, double *accv) {
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
			wx[a] = hatw(hdiv ? d * inv_h : d);
			ax[a] = vk - c0 * d; // v_k + c_k0 * (x_sample - x_p)
		}
#pragma unroll
		for (int b = 0; b < NB; ++b) {
			double d = py - sy[b];
			wy[b] = hatw(hdiv ? d * inv_h : d);
			by[b] = -c1 * d;
		}
#pragma unroll
		for (int c = 0; c < NC; ++c) {
			double d = pz - sz[c];
			wz[c] = hatw(hdiv ? d * inv_h : d);
			cz[c] = -c2 * d;
		}
#pragma unroll
		for (int c = 0; c < NC; ++c) {
#pragma unroll
			for (int b = 0; b < NB; ++b) {
				const double wyz = wy[b] * wz[c], bc = by[b] + cz[c];
#pragma unroll
				for (int a = 0; a < NA; ++a) {
					const double w = wx[a] * wyz;
					const int t = (c * NB + b) * NA + a;
					accw[t] += w;
					accv[t] = fma(w, ax[a] + bc, accv[t]);
				}
			}
		}
	}
}

