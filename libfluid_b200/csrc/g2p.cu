// G1-G4: grid -> particle transfer (reference src/mac_grid.cpp:40-112, src/simulation.cpp:447-560).
//
// The reference evaluates, per velocity component, a trilinear interpolation of 8 face samples and -- for APIC --
// the sum over the same 8 samples of the gradient of the trilinear hat kernel (_calculate_c_vector, :507-521):
// 8 x (3 kernel-gradient components, each a product of three factors and a division by h) per component.  Both are
// the value and the gradient of ONE trilinear function, and the hat weights are separable, so this kernel reduces
// the 8 samples axis by axis (x, then y, then z) carrying "interpolated" and "differenced" partial results:
// ~32 fp64 operations per component instead of ~150, which moves the kernel from the fp64 pipe to the memory
// roofline.  The result differs from the reference's corner-by-corner sum by rounding only (rel. 1e-16 per term;
// the parity tests hold v and c to rel-L2 <= 1e-12).
//
// Compiled with fused multiply-add.
#include "lfk_internal.cuh"

struct G2PArgs {
	const double *px, *py, *pz;
	const double *vs[3];     // particle velocity before the transfer (FLIP only), read through `perm` when given
	double *vd[3];           // particle velocity out
	double *cd[9];           // APIC c rows out
	const double *u, *v, *w;     // face velocities
	const double *uo, *vo, *wo;  // FLIP: the pre-projection snapshot
	const uint32_t *perm;    // FLIP: vs is still in the order before the last sort (NULL: already permuted)
	// multi-GPU: w / wo of the layer BELOW the lower ghost layer (z0 - 2), or NULL.  The position correction can nudge
	// an own particle of the bottom layer into the ghost layer; its z-faces then reach one layer further down.
	const double *w_below, *wo_below;
	double blend;
	unsigned long long *speed2; // max |v|^2 over the particles of this launch (bits of a non-negative double), for cfl()
};

struct FaceFetch { // the 3 clamped cell coordinates per axis of get_face_samples, and their "clamped" bits
	int ci[3][3];
	bool cl[3][3];
};

__device__ __forceinline__ void face_fetch_setup(const GridDesc &G, const long long *gi, FaceFetch &F) {
	const int size[3] = { G.nx, G.ny, G.nz };
#pragma unroll
	for (int a = 0; a < 3; ++a) {
#pragma unroll
		for (int d = 0; d < 3; ++d) {
			long long val = gi[a] + d; // _clamp(val, 1, max) then -1 (src/mac_grid.cpp:42-50)
			if (val < 1) {
				F.ci[a][d] = 0;
				F.cl[a][d] = true;
			} else if (val >= size[a]) {
				F.ci[a][d] = size[a] - 1;
				F.cl[a][d] = true;
			} else {
				F.ci[a][d] = (int)val - 1;
				F.cl[a][d] = false;
			}
		}
	}
}

// the 8 samples of component K: index bit 0 <-> x, bit 1 <-> y, bit 2 <-> z
template <int K> __device__ __forceinline__ void face_samples_comp(const GridDesc &G, const FaceFetch &F,
	const double *__restrict__ comp, const double *__restrict__ below, const int *dsel, double *s) {
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		int bx = k & 1, by = (k >> 1) & 1, bz = (k >> 2) & 1;
		int dx = K == 0 ? bx : dsel[0] + bx;
		int dy = K == 1 ? by : dsel[1] + by;
		int dz = K == 2 ? bz : dsel[2] + bz;
		bool clamped = K == 0 ? F.cl[0][dx] : (K == 1 ? F.cl[1][dy] : F.cl[2][dz]);
		int lz = F.ci[2][dz] - G.z0 + 1;
		// multi-GPU: position correction can nudge a boundary particle into the ghost layer, whose far neighbours are
		// not held by this rank -- the nearest held layer stands in (single GPU: never clamps)
		const double *src = comp;
		if (K == 2 && lz < 0 && below != nullptr) { // the extra layer lives in its own one-layer array
			src = below;
			lz = 0;
		}
		lz = lz < 0 ? 0 : (lz > G.nlz - 1 ? G.nlz - 1 : lz);
		long long idx = F.ci[0][dx] + (long long)G.nx * (F.ci[1][dy] + (long long)G.ny * lz);
		s[k] = clamped ? 0.0 : __ldg(src + idx);
	}
}

// value (and, GRAD, h * gradient) of the trilinear interpolant of the 8 samples at weights (wx, wy, wz) in [0, 1).
// lerp(a, b, t) = a (1 - t) + b t as in the reference (include/fluid/misc.h:20-36); the derivative along an axis is
// s0 * a + b with s0 = -1, or +1 when the weight is exactly 0 (_grad_kernel's `p > 0 ? -1 : 1`, :215-224).
template <bool GRAD> __device__ __forceinline__ void trilinear(const double *s, double wx, double wy, double wz,
	double &val, double *grad) {
	const double ux = 1.0 - wx, uy = 1.0 - wy, uz = 1.0 - wz;
	const double sx0 = wx > 0.0 ? -1.0 : 1.0, sy0 = wy > 0.0 ? -1.0 : 1.0, sz0 = wz > 0.0 ? -1.0 : 1.0;
	double L[4], D[4];
#pragma unroll
	for (int q = 0; q < 4; ++q) { // along x
		const double a = s[2 * q], b = s[2 * q + 1];
		L[q] = fma(b, wx, a * ux);
		if (GRAD) { D[q] = fma(sx0, a, b); }
	}
	double LL[2], DL[2], LD[2];
#pragma unroll
	for (int q = 0; q < 2; ++q) { // along y
		LL[q] = fma(L[2 * q + 1], wy, L[2 * q] * uy);
		if (GRAD) {
			DL[q] = fma(D[2 * q + 1], wy, D[2 * q] * uy);
			LD[q] = fma(sy0, L[2 * q], L[2 * q + 1]);
		}
	}
	val = fma(LL[1], wz, LL[0] * uz); // along z
	if (GRAD) {
		grad[0] = fma(DL[1], wz, DL[0] * uz);
		grad[1] = fma(LD[1], wz, LD[0] * uz);
		grad[2] = fma(sz0, LL[0], LL[1]);
	}
}

// the 8 samples of component K for a particle whose 3 x 3 x 3 cell neighbourhood lies inside the grid (and inside the
// slab's layers): one base index plus constant offsets instead of 24 clamped index computations.  Same addresses, same
// values as face_samples_comp when nothing is clamped.
template <int K> __device__ __forceinline__ void face_samples_interior(const GridDesc &G, const double *__restrict__ comp,
	long long base, const int *dsel, double *s) {
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		const int bx = k & 1, by = (k >> 1) & 1, bz = (k >> 2) & 1;
		const int dx = K == 0 ? bx : dsel[0] + bx;
		const int dy = K == 1 ? by : dsel[1] + by;
		const int dz = K == 2 ? bz : dsel[2] + bz;
		s[k] = __ldg(comp + (base + dx + (long long)G.nx * dy + G.sxy * dz));
	}
}

// Thread per particle.  All 24 face samples of a particle are requested before the first result is stored (the stores
// may alias the sample arrays as far as the compiler can tell, so a component-by-component form serialises four memory
// round trips per particle; the kernel is latency bound), and particles whose 3 x 3 x 3 cell neighbourhood is interior
// take face_samples_interior (index arithmetic was 37 % of the instructions, ncu r1d).  Measured at 256^3 (r2a sweep):
// 4.09 ms against 4.59 ms component by component and 4.47 ms batched with clamped indexing.
template <int METHOD> __device__ __forceinline__ double g2p_one(const GridDesc &G, const G2PArgs &A, unsigned long long i) {
	constexpr bool APIC = METHOD == LFK_METHOD_APIC;
	const double p[3] = { A.px[i], A.py[i], A.pz[i] };
	long long gi[3];
	double t[3], tmid[3];
	int dsel[3];
#pragma unroll
	for (int d = 0; d < 3; ++d) { // compute_cell_index_and_position: no clamping (src/simulation.cpp:13-23)
		double f = div_h(p[d] - G.off[d], G);
		unsigned long long ci = (unsigned long long)f;
		gi[d] = ci > 0x7fffffffull ? 0x7fffffffll : (long long)ci;
		t[d] = f - (double)ci;
		tmid[d] = t[d] - 0.5;
		dsel[d] = 1;
		if (tmid[d] < 0.0) {
			dsel[d] = 0;
			tmid[d] += 1.0;
		}
	}
	FaceFetch F;
	face_fetch_setup(G, gi, F);
	double s[8], vn[3], g[3];
	{
		double s1[8], s2[8];
		// cells gi - 1 .. gi + 1 per axis, none clamped (face_fetch_setup), and their layers held by this rank
		const long long lzlo = gi[2] - 1 - G.z0 + 1;
		const bool interior = gi[0] >= 1 && gi[0] + 2 < G.nx && gi[1] >= 1 && gi[1] + 2 < G.ny && gi[2] >= 1 &&
			gi[2] + 2 < G.nz && lzlo >= 0 && lzlo + 2 <= G.nlz - 1;
		const long long base = (gi[0] - 1) + (long long)G.nx * ((gi[1] - 1) + (long long)G.ny * lzlo);
		if (interior) {
			face_samples_interior<0>(G, A.u, base, dsel, s);
			face_samples_interior<1>(G, A.v, base, dsel, s1);
			face_samples_interior<2>(G, A.w, base, dsel, s2);
		} else {
			face_samples_comp<0>(G, F, A.u, nullptr, dsel, s);
			face_samples_comp<1>(G, F, A.v, nullptr, dsel, s1);
			face_samples_comp<2>(G, F, A.w, A.w_below, dsel, s2);
		}
		double g1[3], g2[3];
		trilinear<APIC>(s, t[0], tmid[1], tmid[2], vn[0], g);
		trilinear<APIC>(s1, tmid[0], t[1], tmid[2], vn[1], g1);
		trilinear<APIC>(s2, tmid[0], tmid[1], t[2], vn[2], g2);
		if (METHOD == LFK_METHOD_FLIP) { // v = v_new + (v_p - v_old) * blend (:463-505)
			double vold[3], dummy[3];
			if (interior) {
				face_samples_interior<0>(G, A.uo, base, dsel, s);
				face_samples_interior<1>(G, A.vo, base, dsel, s1);
				face_samples_interior<2>(G, A.wo, base, dsel, s2);
			} else {
				face_samples_comp<0>(G, F, A.uo, nullptr, dsel, s);
				face_samples_comp<1>(G, F, A.vo, nullptr, dsel, s1);
				face_samples_comp<2>(G, F, A.wo, A.wo_below, dsel, s2);
			}
			const unsigned long long src = A.perm ? (unsigned long long)A.perm[i] : i;
			const double vp[3] = { A.vs[0][src], A.vs[1][src], A.vs[2][src] };
			trilinear<false>(s, t[0], tmid[1], tmid[2], vold[0], dummy);
			trilinear<false>(s1, tmid[0], t[1], tmid[2], vold[1], dummy);
			trilinear<false>(s2, tmid[0], tmid[1], t[2], vold[2], dummy);
#pragma unroll
			for (int d = 0; d < 3; ++d) {
				vn[d] = vn[d] + (vp[d] - vold[d]) * A.blend;
			}
		}
		if (APIC) {
#pragma unroll
			for (int k = 0; k < 3; ++k) {
				A.cd[k][i] = div_h(g[k], G);
				A.cd[3 + k][i] = div_h(g1[k], G);
				A.cd[6 + k][i] = div_h(g2[k], G);
			}
		}
		A.vd[0][i] = vn[0];
		A.vd[1][i] = vn[1];
		A.vd[2][i] = vn[2];
	}
	// squared_length() as the reference's cfl() evaluates it (src/simulation.cpp:199-205): no contraction
	return __dadd_rn(__dadd_rn(__dmul_rn(vn[0], vn[0]), __dmul_rn(vn[1], vn[1])), __dmul_rn(vn[2], vn[2]));
}

// The new velocities' largest |v|^2 is folded in here, so that the cfl() of the next step needs no pass of its own over
// the particles (3.1 GB at 256^3, 0.48 ms).  The maximum does not depend on the order; a NaN speed is skipped like
// std::max(m, s) skips it.
#ifndef G2P_THREADS
#define G2P_THREADS 128
#endif
#ifndef G2P_BLOCKS
#define G2P_BLOCKS 7 // resident blocks per SM the register budget of the PIC / APIC instantiations is set for (72 registers)
#endif
template <int METHOD> __global__ void __launch_bounds__(G2P_THREADS, METHOD == LFK_METHOD_FLIP ? 1 : G2P_BLOCKS) k_g2p(GridDesc G, G2PArgs A, unsigned long long n) {
	const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	double s2 = 0.0;
	if (i < n) {
		const double s = g2p_one<METHOD>(G, A, i);
		s2 = 0.0 < s ? s : 0.0;
	}
	// non-negative doubles order like their bit patterns: maximum of the high words, then of the low words among the
	// lanes that hold it
	const unsigned long long bits = (unsigned long long)__double_as_longlong(s2);
	const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
	const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
	const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
	if ((threadIdx.x & 31) == 0) {
		const unsigned long long m = ((unsigned long long)mhi << 32) | mlo;
		if (m > *(volatile unsigned long long*)A.speed2) { atomicMax(A.speed2, m); }
	}
}

int lfkp_g2p(lfk_ctx *c) {
	PhaseTimer T(c, LFK_PHASE_G2P);
	const int method = c->prm.method;
	// While the v / c payload of a lean sort is still pending: PIC / APIC overwrite v (APIC: and c) without reading
	// them, so the pending permutation is simply dropped.  FLIP reads the old particle velocity through the
	// permutation and writes the result to the alternate buffers, which then become current.
	if (c->c_deferred && method != LFK_METHOD_APIC) { LFK_TRY(lfkp_permute_c(c)); }
	const bool indirect = c->v_deferred && method == LFK_METHOD_FLIP;
	if (c->nranks > 1) { // the samples reach into the ghost layers and, for z-faces, one layer below the lower one
		for (int d = 0; d < 3; ++d) { LFK_TRY(lfkx_halo_f64(c, c->vel[d])); }
		LFK_TRY(lfkx_layer_below(c, c->vel[2], c->wlow[0]));
	}
	// the launch also leaves max |v|^2 of the own particles in the cfl cache (lfkp_cfl)
	unsigned long long *speed2 = (unsigned long long*)(c->d_reduce + LFK_REDUCE_SPEED2);
	LFK_CUDA(c, cudaMemsetAsync(speed2, 0, sizeof(double), c->stream));
	if (c->np > 0) {
		G2PArgs A;
		A.speed2 = speed2;
		// own particles are entries [first, first + np); the permutation holds absolute source indices
		const uint64_t o = c->first;
		A.px = c->P.f[PF_PX] + o;
		A.py = c->P.f[PF_PY] + o;
		A.pz = c->P.f[PF_PZ] + o;
		for (int d = 0; d < 3; ++d) {
			A.vs[d] = indirect ? c->P.f[PF_VX + d] : c->P.f[PF_VX + d] + o;
			A.vd[d] = (indirect ? c->Palt.f[PF_VX + d] : c->P.f[PF_VX + d]) + o;
		}
		for (int k = 0; k < 9; ++k) { A.cd[k] = c->P.f[PF_C0 + k] + o; }
		A.u = c->vel[0]; A.v = c->vel[1]; A.w = c->vel[2];
		A.uo = c->vel_old[0]; A.vo = c->vel_old[1]; A.wo = c->vel_old[2];
		A.perm = indirect ? c->perm + o : nullptr;
		A.w_below = c->rank > 0 ? c->wlow[0] : nullptr;
		A.wo_below = c->rank > 0 ? c->wlow[1] : nullptr;
		A.blend = c->prm.blending_factor;
		unsigned nb = lfk_blocks((long long)c->np, G2P_THREADS);
		unsigned long long n = c->np;
		switch (method) {
		case LFK_METHOD_PIC:
			LFK_LAUNCH(c, k_g2p<LFK_METHOD_PIC>, nb, G2P_THREADS, 0, c->g, A, n);
			break;
		case LFK_METHOD_FLIP:
			LFK_LAUNCH(c, k_g2p<LFK_METHOD_FLIP>, nb, G2P_THREADS, 0, c->g, A, n);
			break;
		default:
			LFK_LAUNCH(c, k_g2p<LFK_METHOD_APIC>, nb, G2P_THREADS, 0, c->g, A, n);
			break;
		}
		if (indirect) {
			for (int d = 0; d < 3; ++d) {
				double *t = c->P.f[PF_VX + d]; c->P.f[PF_VX + d] = c->Palt.f[PF_VX + d]; c->Palt.f[PF_VX + d] = t;
			}
		}
	}
	c->v_deferred = false;
	c->c_deferred = false;
	c->speed2_valid = true;
	return 0;
}
