// P2G: particle -> MAC-face transfer (PIC / FLIP / APIC), fused with normalisation, cell classification,
// boundary-face zeroing, the FLIP snapshot and (optionally) the gravity update.
// Reference: simulation::_transfer_to_grid_{pic,flip,apic}, src/simulation.cpp:293-412, 428-445, 72-78.
//
// Formulation: cell-sorted GATHER.  The reference itself gathers per cell over the 27 neighbouring cells of the
// sorted particle table; because particles are sorted by raw cell index, the cells x-1..x+1 of one (y, z) row form
// ONE contiguous particle range, so a cell reads 9 contiguous ranges.  There are no atomics and the summation
// order is fixed (rows in z, y order, particles in stable-sorted order) => bit-reproducible, and identical to the
// order of the CPU restatement.
#include "lfk_internal.cuh"

#include <cstdlib>

struct P2GParams {
	double h, half;
	double gdt[3];
	int method;
	int add_gravity;
};

// per-axis linear hat weight max(0, 1 - |d|)
__device__ __forceinline__ double hat(double d) {
	return dmax_std(0.0, 1.0 - fabs(d));
}

template <int METHOD> __global__ void __launch_bounds__(128) k_p2g_gather(GridDesc G, P2GParams Q, ParticleSoA P,
	const uint32_t *__restrict__ begin, const double *__restrict__ cxs, const double *__restrict__ cys,
	const double *__restrict__ czs, double *__restrict__ u, double *__restrict__ v, double *__restrict__ w,
	double *__restrict__ uo, double *__restrict__ vo, double *__restrict__ wo, uint8_t *__restrict__ typ) {
	long long own = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (own >= G.nown) { return; }
	int x = (int)(own % G.nx);
	long long rest = own / G.nx;
	int y = (int)(rest % G.ny);
	int lz = (int)(rest / G.ny) + 1;
	int z = lz - 1 + G.z0;
	long long me = own + G.sxy;

	// cell centre (reference: repeated addition of cell_size, see lfk_api.cu) and the three +face positions
	const double xc = cxs[x], yc = cys[y], zc = czs[z];
	const double xf = xc + Q.half, yf = yc + Q.half, zf = zc + Q.half;
	const double *__restrict__ px = P.f[PF_PX], *__restrict__ py = P.f[PF_PY], *__restrict__ pz = P.f[PF_PZ];

	double sw0 = 0.0, sw1 = 0.0, sw2 = 0.0, sv0 = 0.0, sv1 = 0.0, sv2 = 0.0;
	int x0 = x < 1 ? 0 : x - 1, x1 = x + 2 < G.nx ? x + 2 : G.nx;
	int y0 = y < 1 ? 0 : y - 1, y1 = y + 2 < G.ny ? y + 2 : G.ny;
	for (int dz = -1; dz <= 1; ++dz) { // the z ghost layers are empty at the domain boundary
		int cz = z + dz;
		if (cz < 0 || cz >= G.nz) { continue; }
		for (int cy = y0; cy < y1; ++cy) {
			long long row = (long long)G.nx * (cy + (long long)G.ny * (lz + dz));
			uint32_t qb = begin[row + x0], qe = begin[row + x1];
			for (uint32_t q = qb; q < qe; ++q) {
				double ppx = px[q], ppy = py[q], ppz = pz[q];
				double dxc = ppx - xc, dyc = ppy - yc, dzc = ppz - zc;
				double dxf = ppx - xf, dyf = ppy - yf, dzf = ppz - zf;
				double w0, w1, w2;
				if (METHOD == LFK_METHOD_APIC) { // weights WITHOUT /h (src/simulation.cpp:367-369)
					double hxc = hat(dxc), hyc = hat(dyc), hzc = hat(dzc);
					w0 = hat(dxf) * hyc * hzc;
					w1 = hxc * hat(dyf) * hzc;
					w2 = hxc * hyc * hat(dzf);
				} else { // PIC / FLIP divide (src/simulation.cpp:313-315)
					double hxc = hat(div_h(dxc, G)), hyc = hat(div_h(dyc, G)), hzc = hat(div_h(dzc, G));
					w0 = hat(div_h(dxf, G)) * hyc * hzc;
					w1 = hxc * hat(div_h(dyf, G)) * hzc;
					w2 = hxc * hyc * hat(div_h(dzf, G));
				}
				if (w0 == 0.0 && w1 == 0.0 && w2 == 0.0) { continue; } // adds exact zeros in the reference
				double v0 = P.f[PF_VX][q], v1 = P.f[PF_VY][q], v2 = P.f[PF_VZ][q];
				if (METHOD == LFK_METHOD_APIC) { // affine term c_k . (x_face_k - x_p)  (:371-375)
					double a0 = 0.0, a1 = 0.0, a2 = 0.0;
					a0 += P.f[PF_C0 + 0][q] * (xf - ppx);
					a0 += P.f[PF_C0 + 1][q] * (yc - ppy);
					a0 += P.f[PF_C0 + 2][q] * (zc - ppz);
					a1 += P.f[PF_C0 + 3][q] * (xc - ppx);
					a1 += P.f[PF_C0 + 4][q] * (yf - ppy);
					a1 += P.f[PF_C0 + 5][q] * (zc - ppz);
					a2 += P.f[PF_C0 + 6][q] * (xc - ppx);
					a2 += P.f[PF_C0 + 7][q] * (yc - ppy);
					a2 += P.f[PF_C0 + 8][q] * (zf - ppz);
					v0 += a0;
					v1 += a1;
					v2 += a2;
				}
				sw0 += w0;
				sw1 += w1;
				sw2 += w2;
				sv0 += w0 * v0;
				sv1 += w1 * v1;
				sv2 += w2 * v2;
			}
		}
	}
	double r0 = sw0 > 1e-6 ? sv0 / sw0 : 0.0; // src/simulation.cpp:380-386
	double r1 = sw1 > 1e-6 ? sv1 / sw1 : 0.0;
	double r2 = sw2 > 1e-6 ? sv2 / sw2 : 0.0;
	// classification (:388-393)
	uint8_t t = typ[me];
	if (t != LFK_CELL_SOLID) {
		t = (begin[me + 1] - begin[me]) > 0 ? LFK_CELL_FLUID : LFK_CELL_AIR;
		typ[me] = t;
	}
	const bool bx = x == G.nx - 1, by = y == G.ny - 1, bz = z == G.nz - 1;
	if (METHOD == LFK_METHOD_FLIP) { // _old_grid = _grid, boundary faces zeroed on the copy only (:340-344)
		uo[me] = bx ? 0.0 : r0;
		vo[me] = by ? 0.0 : r1;
		wo[me] = bz ? 0.0 : r2;
	}
	if (METHOD == LFK_METHOD_APIC) { // _remove_boundary_velocities(_grid) (:397)
		if (bx) { r0 = 0.0; }
		if (by) { r1 = 0.0; }
		if (bz) { r2 = 0.0; }
	}
	if (Q.add_gravity) { // every cell, every component (:72-78)
		r0 += Q.gdt[0];
		r1 += Q.gdt[1];
		r2 += Q.gdt[2];
	}
	u[me] = r0;
	v[me] = r1;
	w[me] = r2;
}

__global__ void k_gravity(GridDesc G, double g0, double g1, double g2, double *__restrict__ u,
	double *__restrict__ v, double *__restrict__ w) {
	long long own = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (own >= G.nown) { return; }
	long long me = own + G.sxy;
	u[me] += g0;
	v[me] += g1;
	w[me] += g2;
}

int lfkg_p2g_march(lfk_ctx *c, double gravity_dt, bool add_gravity); // p2g_march.cu

int lfkg_p2g(lfk_ctx *c, double gravity_dt, bool add_gravity) {
	PhaseTimer T(c, LFK_PHASE_P2G);
	LFK_REQUIRE(c, c->table_valid, LFK_E_STATE, "lfk_p2g needs the cell table of lfk_hash");
	const GridDesc &G = c->g;
	// The simple gather kernel below is the A/B reference -- and the production path for one corner of the
	// reference's behaviour: APIC evaluates its hat weights on the raw distance, not distance / h
	// (src/simulation.cpp:367-369), so for h < 1 a particle still weighs on the +face of the cell BEHIND its
	// neighbour along the staggered axis (|d| up to 2 h < 1 where h < 0.5, up to 1 otherwise), which the reference's
	// 27-cell gather visits.  The scatter kernels give every cell 2 faces along the staggered axis (all that can be
	// reached when the weight's support is one cell), so that case goes to the gather kernel, which walks exactly the
	// reference's 27 cells.  h >= 1 (every shipped configuration has h == 1) is unaffected.
	const bool wide_apic = c->prm.method == LFK_METHOD_APIC && G.h < 1.0;
	const bool use_gather = c->tune.p2g == LFK_TUNE_P2G_GATHER || wide_apic;
	if (!use_gather) {
		LFK_TRY(lfkg_p2g_march(c, gravity_dt, add_gravity));
		if (c->nranks > 1 && c->prm.method == LFK_METHOD_FLIP) { // FLIP's G2P samples the snapshot in the ghost layers
			for (int d = 0; d < 3; ++d) { LFK_TRY(lfkx_halo_f64(c, c->vel_old[d])); }
			LFK_TRY(lfkx_layer_below(c, c->vel_old[2], c->wlow[1]));
		}
		c->system_valid = false;
		c->pressure_valid = false;
		return 0;
	}
	LFK_TRY(lfkp_materialise_vc(c));
	P2GParams Q;
	Q.h = G.h;
	Q.half = 0.5 * G.h;
	Q.method = c->prm.method;
	Q.add_gravity = add_gravity ? 1 : 0;
	for (int d = 0; d < 3; ++d) {
		Q.gdt[d] = c->prm.gravity[d] * gravity_dt;
	}
	unsigned nb = lfk_blocks(G.nown, 128);
	switch (c->prm.method) {
	case LFK_METHOD_PIC:
		LFK_LAUNCH(c, k_p2g_gather<LFK_METHOD_PIC>, nb, 128, 0, G, Q, c->P, c->begin, c->ctr[0], c->ctr[1], c->ctr[2],
			c->vel[0], c->vel[1], c->vel[2], c->vel_old[0], c->vel_old[1], c->vel_old[2], c->typ);
		break;
	case LFK_METHOD_FLIP:
		LFK_LAUNCH(c, k_p2g_gather<LFK_METHOD_FLIP>, nb, 128, 0, G, Q, c->P, c->begin, c->ctr[0], c->ctr[1], c->ctr[2],
			c->vel[0], c->vel[1], c->vel[2], c->vel_old[0], c->vel_old[1], c->vel_old[2], c->typ);
		break;
	default:
		LFK_LAUNCH(c, k_p2g_gather<LFK_METHOD_APIC>, nb, 128, 0, G, Q, c->P, c->begin, c->ctr[0], c->ctr[1], c->ctr[2],
			c->vel[0], c->vel[1], c->vel[2], c->vel_old[0], c->vel_old[1], c->vel_old[2], c->typ);
		break;
	}
	if (c->nranks > 1 && c->prm.method == LFK_METHOD_FLIP) { // as above
		for (int d = 0; d < 3; ++d) { LFK_TRY(lfkx_halo_f64(c, c->vel_old[d])); }
		LFK_TRY(lfkx_layer_below(c, c->vel_old[2], c->wlow[1]));
	}
	c->system_valid = false;
	c->pressure_valid = false;
	return 0;
}

int lfkg_gravity(lfk_ctx *c, double dt) {
	PhaseTimer T(c, LFK_PHASE_P2G);
	const GridDesc &G = c->g;
	LFK_LAUNCH(c, k_gravity, lfk_blocks(G.nown, 256), 256, 0, G, c->prm.gravity[0] * dt, c->prm.gravity[1] * dt,
		c->prm.gravity[2] * dt, c->vel[0], c->vel[1], c->vel[2]);
	c->system_valid = false;
	return 0;
}
