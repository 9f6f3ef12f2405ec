/* TEST INFRASTRUCTURE ONLY -- see fluid_oracle.h.  Plain-C restatement of the libfluid hot path.
 * Compiled with -ffp-contract=off so that, like the reference's x86-64 build, no FMA contraction happens and
 * the results can be pinned bit-for-bit against oracle/_ref. */
#define _POSIX_C_SOURCE 200809L
#include "fluid_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define RAW(P, x, y, z) ((uint64_t)(x) + (P)->nx * ((uint64_t)(y) + (P)->ny * (uint64_t)(z)))

static inline double dmax(double a, double b) { /* std::max(a, b) */
	return (a < b) ? b : a;
}
static inline double dclamp(double v, double lo, double hi) { /* std::clamp */
	return (v < lo) ? lo : (hi < v) ? hi : v;
}
static inline double dot3(const double *a, const double *b) { /* vec_ops::dot, include/fluid/math/vec.h:108-121 */
	double r = 0.0;
	r += a[0] * b[0];
	r += a[1] * b[1];
	r += a[2] * b[2];
	return r;
}
static inline uint64_t trunc_index(double v) { /* static_cast<std::size_t>(double) */
	return (uint64_t)v;
}

/* ------------------------------------------------------------------------------------------------ K1 */
void fo_cell_keys(const fo_params *P, size_t n, const double *pos, uint64_t *key) {
	const uint64_t size[3] = { P->nx, P->ny, P->nz };
	for (size_t i = 0; i < n; ++i) {
		uint64_t idx[3];
		for (int d = 0; d < 3; ++d) {
			double g = (pos[3 * i + d] - P->off[d]) / P->h; /* src/simulation.cpp:253 */
			uint64_t v = trunc_index(dmax(g, 0.0));        /* :256 */
			idx[d] = v < size[d] - 1 ? v : size[d] - 1;
		}
		key[i] = RAW(P, idx[0], idx[1], idx[2]);
	}
}

/* ------------------------------------------------------------------------------------------------ K2 */
size_t fo_cell_ranges(const fo_params *P, size_t n, const uint64_t *k, uint64_t *begin, uint64_t *count,
	uint64_t *fluid_cells) {
	size_t nc = (size_t)(P->nx * P->ny * P->nz), nf = 0;
	memset(begin, 0, nc * sizeof(uint64_t)); /* reset_space_hash, src/simulation.cpp:131-134 */
	memset(count, 0, nc * sizeof(uint64_t));
	if (n == 0) {
		return 0;
	}
	uint64_t last = k[0], cnt = 1; /* src/simulation.cpp:273-290; begin of the first run stays 0 */
	fluid_cells[nf++] = last;
	for (size_t i = 1; i < n; ++i, ++cnt) {
		uint64_t cur = k[i];
		if (cur != last) {
			count[last] = cnt;
			cnt = 0;
			fluid_cells[nf++] = cur;
			begin[cur] = i;
			last = cur;
		}
	}
	count[last] = cnt;
	return nf;
}

size_t fo_hash(const fo_params *P, size_t n, const uint64_t *key, uint64_t *perm, uint64_t *begin, uint64_t *count,
	uint64_t *fluid_cells) {
	size_t nc = (size_t)(P->nx * P->ny * P->nz);
	uint64_t *cursor = (uint64_t*)calloc(nc + 1, sizeof(uint64_t));
	for (size_t i = 0; i < n; ++i) {
		++cursor[key[i] + 1];
	}
	for (size_t c = 0; c < nc; ++c) {
		cursor[c + 1] += cursor[c];
	}
	uint64_t *sorted = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
	for (size_t i = 0; i < n; ++i) { /* stable */
		uint64_t d = cursor[key[i]]++;
		perm[d] = i;
		sorted[d] = key[i];
	}
	size_t nf = fo_cell_ranges(P, n, sorted, begin, count, fluid_cells);
	free(sorted);
	free(cursor);
	return nf;
}

/* ------------------------------------------------------------------------------------------- P1 - P3 */
static inline double kernel3(const double *d) { /* _kernel, src/simulation.cpp:207-213 */
	return dmax(0.0, 1.0 - fabs(d[0])) * dmax(0.0, 1.0 - fabs(d[1])) * dmax(0.0, 1.0 - fabs(d[2]));
}

static void remove_boundary_velocities(const fo_params *P, double *gvel) { /* src/simulation.cpp:428-445 */
	if (P->nx * P->ny * P->nz == 0) {
		return;
	}
	for (uint64_t z = 0; z < P->nz; ++z) {
		for (uint64_t y = 0; y < P->ny; ++y) {
			gvel[3 * RAW(P, P->nx - 1, y, z) + 0] = 0.0;
		}
		for (uint64_t x = 0; x < P->nx; ++x) {
			gvel[3 * RAW(P, x, P->ny - 1, z) + 1] = 0.0;
		}
	}
	for (uint64_t y = 0; y < P->ny; ++y) {
		for (uint64_t x = 0; x < P->nx; ++x) {
			gvel[3 * RAW(P, x, y, P->nz - 1) + 2] = 0.0;
		}
	}
}

void fo_p2g(const fo_params *P, size_t n, const double *pos, const double *vel, const double *c,
	const uint64_t *begin, const uint64_t *count, double *gvel, uint8_t *type, double *old_gvel) {
	(void)n;
	const int apic = P->method == FO_APIC;
	const double half = 0.5 * P->h;
	double zpos = P->off[2] + half; /* face positions by repeated addition, src/simulation.cpp:294-300,347-353 */
	for (uint64_t z = 0; z < P->nz; ++z, zpos += P->h) {
		double zface = zpos + half, ypos = P->off[1] + half;
		for (uint64_t y = 0; y < P->ny; ++y, ypos += P->h) {
			double yface = ypos + half, xpos = P->off[0] + half;
			for (uint64_t x = 0; x < P->nx; ++x, xpos += P->h) {
				double xface = xpos + half;
				const double face[3][3] = { { xface, ypos, zpos }, { xpos, yface, zpos }, { xpos, ypos, zface } };
				double sum_vel[3] = { 0.0, 0.0, 0.0 }, sum_w[3] = { 0.0, 0.0, 0.0 };
				/* _for_all_nearby_particles(center, (1,1,1), (1,1,1)): include/fluid/simulation.h:212-223,
				 * include/fluid/data_structures/grid.h:117-135 -- z outermost, x innermost */
				uint64_t x0 = x < 1 ? 0 : x - 1, y0 = y < 1 ? 0 : y - 1, z0 = z < 1 ? 0 : z - 1;
				uint64_t x1 = x + 2 < P->nx ? x + 2 : P->nx, y1 = y + 2 < P->ny ? y + 2 : P->ny;
				uint64_t z1 = z + 2 < P->nz ? z + 2 : P->nz;
				for (uint64_t cz = z0; cz < z1; ++cz) {
					for (uint64_t cy = y0; cy < y1; ++cy) {
						for (uint64_t cx = x0; cx < x1; ++cx) {
							uint64_t cell = RAW(P, cx, cy, cz);
							for (uint64_t q = begin[cell], k = 0; k < count[cell]; ++q, ++k) {
								const double *pp = pos + 3 * q, *pv = vel + 3 * q;
								double w[3], aff[3] = { 0.0, 0.0, 0.0 };
								for (int d = 0; d < 3; ++d) {
									double diff[3] = { pp[0] - face[d][0], pp[1] - face[d][1], pp[2] - face[d][2] };
									if (apic) { /* weights WITHOUT /h, src/simulation.cpp:367-369 */
										w[d] = kernel3(diff);
										double fm[3] = { face[d][0] - pp[0], face[d][1] - pp[1], face[d][2] - pp[2] };
										aff[d] = dot3(c + 9 * q + 3 * d, fm); /* :372-374 */
									} else {    /* PIC divides, :313-315 */
										diff[0] /= P->h;
										diff[1] /= P->h;
										diff[2] /= P->h;
										w[d] = kernel3(diff);
									}
								}
								for (int d = 0; d < 3; ++d) {
									sum_w[d] += w[d];
									sum_vel[d] += w[d] * (apic ? pv[d] + aff[d] : pv[d]);
								}
							}
						}
					}
				}
				uint64_t me = RAW(P, x, y, z);
				for (int d = 0; d < 3; ++d) { /* :380-386 */
					gvel[3 * me + d] = sum_w[d] > 1e-6 ? sum_vel[d] / sum_w[d] : 0.0;
				}
				if (type[me] != FO_SOLID) { /* :388-393 */
					type[me] = count[me] > 0 ? FO_FLUID : FO_AIR;
				}
			}
		}
	}
	size_t nc = (size_t)(P->nx * P->ny * P->nz);
	if (P->method == FO_FLIP) { /* :340-344 */
		memcpy(old_gvel, gvel, nc * 3 * sizeof(double));
		remove_boundary_velocities(P, old_gvel);
	} else if (apic) { /* :397 */
		remove_boundary_velocities(P, gvel);
	}
}

/* ------------------------------------------------------------------------------------------------ G0 */
void fo_gravity(const fo_params *P, double dt, double *gvel) {
	size_t nc = (size_t)(P->nx * P->ny * P->nz);
	double gdt[3] = { P->g[0] * dt, P->g[1] * dt, P->g[2] * dt }; /* gravity * dt, then += */
	for (size_t i = 0; i < nc; ++i) {
		gvel[3 * i + 0] += gdt[0];
		gvel[3 * i + 1] += gdt[1];
		gvel[3 * i + 2] += gdt[2];
	}
}

/* ------------------------------------------------------------------------------------------- S1 - S3 */
static inline void from_raw(const fo_params *P, uint64_t raw, uint64_t *xyz) { /* grid.h:225-236 */
	xyz[0] = raw % P->nx;
	raw /= P->nx;
	xyz[1] = raw % P->ny;
	raw /= P->ny;
	xyz[2] = raw;
}
/* mac_grid::get_cell_and_type (src/mac_grid.cpp:10-38): unsigned wrap-around makes "negative" indices out of
 * range; out of range reads as solid.  Returns the type, *inside tells whether the cell exists. */
static inline uint8_t type_at(const fo_params *P, const uint8_t *type, uint64_t x, uint64_t y, uint64_t z, int *inside) {
	if (x >= P->nx || y >= P->ny || z >= P->nz) {
		*inside = 0;
		return FO_SOLID;
	}
	*inside = 1;
	return type[RAW(P, x, y, z)];
}

void fo_solver_setup(const fo_params *P, const double *gvel, const uint8_t *type, size_t nf,
	const uint64_t *fluid_cells, uint64_t *index_map, uint8_t *flags, double *b) {
	size_t nc = (size_t)(P->nx * P->ny * P->nz);
	for (size_t i = 0; i < nc; ++i) {
		index_map[i] = FO_NOT_FLUID;
	}
	for (size_t i = 0; i < nf; ++i) { /* src/pressure_solver.cpp:150-155 */
		index_map[fluid_cells[i]] = i;
	}
	static const int off6[6][3] = { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 }, { -1, 0, 0 }, { 0, -1, 0 }, { 0, 0, -1 } };
	double scale = 1.0 / P->h;
	for (size_t i = 0; i < nf; ++i) {
		uint64_t p[3];
		int in;
		from_raw(P, fluid_cells[i], p);
		unsigned nonsolid = 0; /* :160-178 */
		for (int k = 0; k < 6; ++k) {
			nonsolid += type_at(P, type, p[0] + off6[k][0], p[1] + off6[k][1], p[2] + off6[k][2], &in) != FO_SOLID;
		}
		unsigned fx = type_at(P, type, p[0] + 1, p[1], p[2], &in) == FO_FLUID;
		unsigned fy = type_at(P, type, p[0], p[1] + 1, p[2], &in) == FO_FLUID;
		unsigned fz = type_at(P, type, p[0], p[1], p[2] + 1, &in) == FO_FLUID;
		flags[i] = (uint8_t)(nonsolid | fx << 3 | fy << 4 | fz << 5);

		const double *v = gvel + 3 * fluid_cells[i]; /* :180-242 */
		double value = -(v[0] + v[1] + v[2]);
		for (int d = 0; d < 3; ++d) {
			if (p[d] > 0) {
				uint64_t q[3] = { p[0], p[1], p[2] };
				--q[d];
				uint64_t nb = RAW(P, q[0], q[1], q[2]);
				value += gvel[3 * nb + d];
				if (type[nb] == FO_SOLID) {
					value -= gvel[3 * nb + d];
				}
			}
		}
		for (int d = 0; d < 3; ++d) {
			uint64_t q[3] = { p[0], p[1], p[2] };
			++q[d];
			if (type_at(P, type, q[0], q[1], q[2], &in) == FO_SOLID) {
				value += v[d];
			}
		}
		b[i] = scale * value;
	}
}

/* neighbour ordinal helpers, include/fluid/pressure_solver.h:59-71 */
static inline uint64_t neg_index(const fo_params *P, const uint64_t *map, const uint64_t *p, int d) {
	if (p[d] > 0) {
		uint64_t q[3] = { p[0], p[1], p[2] };
		--q[d];
		return map[RAW(P, q[0], q[1], q[2])];
	}
	return FO_NOT_FLUID;
}
static inline uint64_t pos_index(const fo_params *P, const uint64_t *map, const uint64_t *p, int d) {
	const uint64_t size[3] = { P->nx, P->ny, P->nz };
	if (p[d] + 1 < size[d]) {
		uint64_t q[3] = { p[0], p[1], p[2] };
		++q[d];
		return map[RAW(P, q[0], q[1], q[2])];
	}
	return FO_NOT_FLUID;
}
#define FLAG_N(f) ((double)((f) & 7u))
#define FLAG_POS(f, d) ((double)(((f) >> (3 + (d))) & 1u))

void fo_apply_a(const fo_params *P, double a_scale, size_t nf, const uint64_t *fluid_cells,
	const uint64_t *map, const uint8_t *flags, const double *v, double *out) {
	for (size_t i = 0; i < nf; ++i) { /* src/pressure_solver.cpp:334-362 */
		uint64_t p[3];
		from_raw(P, fluid_cells[i], p);
		double value = FLAG_N(flags[i]) * v[i];
		for (int d = 0; d < 3; ++d) {
			uint64_t j = neg_index(P, map, p, d);
			if (j != FO_NOT_FLUID) {
				value -= FLAG_POS(flags[j], d) * v[j];
			}
		}
		for (int d = 0; d < 3; ++d) {
			uint64_t j = pos_index(P, map, p, d);
			if (j != FO_NOT_FLUID) {
				value -= FLAG_POS(flags[i], d) * v[j];
			}
		}
		out[i] = a_scale * value;
	}
}

static void mic0_setup(const fo_params *P, double a_scale, double tau, double sigma, size_t nf,
	const uint64_t *fluid_cells, const uint64_t *map, const uint8_t *flags, double *precon) {
	for (size_t i = 0; i < nf; ++i) { /* src/pressure_solver.cpp:244-294 */
		uint64_t p[3];
		from_raw(P, fluid_cells[i], p);
		double neg_e = 0.0, neg_e_tau = 0.0;
		for (int d = 0; d < 3; ++d) {
			uint64_t j = neg_index(P, map, p, d);
			if (j != FO_NOT_FLUID) {
				int o1 = (d + 1) % 3, o2 = (d + 2) % 3;
				if (d == 1) { /* y: (xpos + zpos); keep the reference's operand order */
					o1 = 0;
					o2 = 2;
				} else if (d == 2) { /* z: (xpos + ypos) */
					o1 = 0;
					o2 = 1;
				}
				double pj = precon[j], ap = FLAG_POS(flags[j], d) * pj;
				neg_e += ap * ap;
				neg_e_tau += (FLAG_POS(flags[j], d) * (FLAG_POS(flags[j], o1) + FLAG_POS(flags[j], o2))) * pj * pj;
			}
		}
		double nn = FLAG_N(flags[i]);
		double e = nn - (neg_e + tau * neg_e_tau) * a_scale;
		if (e < sigma * nn) {
			e = nn;
		}
		precon[i] = 1.0 / sqrt(e * a_scale);
	}
}

static void mic0_apply(const fo_params *P, double a_scale, size_t nf, const uint64_t *fluid_cells,
	const uint64_t *map, const uint8_t *flags, const double *precon, const double *r, double *q, double *z) {
	for (size_t i = 0; i < nf; ++i) { /* L q = r, src/pressure_solver.cpp:300-314 */
		uint64_t p[3];
		from_raw(P, fluid_cells[i], p);
		double neg_t = 0.0;
		for (int d = 0; d < 3; ++d) {
			uint64_t j = neg_index(P, map, p, d);
			if (j != FO_NOT_FLUID) {
				neg_t += FLAG_POS(flags[j], d) * precon[j] * q[j];
			}
		}
		q[i] = (r[i] + a_scale * neg_t) * precon[i];
	}
	for (size_t i = nf; i > 0; ) { /* L^T z = q, :315-331 */
		--i;
		uint64_t p[3];
		from_raw(P, fluid_cells[i], p);
		double neg_t = 0.0;
		for (int d = 0; d < 3; ++d) {
			uint64_t j = pos_index(P, map, p, d);
			if (j != FO_NOT_FLUID) {
				neg_t += FLAG_POS(flags[i], d) * z[j];
			}
		}
		z[i] = (q[i] + a_scale * precon[i] * neg_t) * precon[i];
	}
}

static double dyn_dot(size_t n, const double *a, const double *b) { /* vec.h:162-171 */
	double r = 0.0;
	for (size_t i = 0; i < n; ++i) {
		r += a[i] * b[i];
	}
	return r;
}
static void muladd(size_t n, double *out, const double *a, const double *b, double s) { /* :364-370 */
	for (size_t i = 0; i < n; ++i) {
		out[i] = a[i] + s * b[i];
	}
}

size_t fo_solve(const fo_params *P, double dt, size_t nf, const uint64_t *fluid_cells, const uint64_t *map,
	const uint8_t *flags, const double *b, double tau, double sigma, double tolerance, size_t max_iterations,
	double *p, double *residual) {
	double a_scale = dt / (P->rho * P->h * P->h); /* src/pressure_solver.cpp:22 */
	size_t alloc = nf ? nf : 1;
	double *precon = (double*)calloc(alloc, sizeof(double));
	mic0_setup(P, a_scale, tau, sigma, nf, fluid_cells, map, flags, precon);
	*residual = 0.0;
	memset(p, 0, nf * sizeof(double));
	double tot = 0.0;
	for (size_t i = 0; i < nf; ++i) {
		tot += b[i] * b[i];
	}
	if (tot < 1e-6) { /* :29-35 */
		free(precon);
		return 0;
	}
	double *r = (double*)malloc(alloc * sizeof(double));
	double *z = (double*)calloc(alloc, sizeof(double)), *q = (double*)calloc(alloc, sizeof(double));
	double *s = (double*)malloc(alloc * sizeof(double));
	memcpy(r, b, nf * sizeof(double));
	mic0_apply(P, a_scale, nf, fluid_cells, map, flags, precon, r, q, z);
	memcpy(s, z, nf * sizeof(double));
	double sigma_ps = dyn_dot(nf, z, r);
	size_t i = 0;
	for (; i < max_iterations; ++i) { /* :44-69 */
		fo_apply_a(P, a_scale, nf, fluid_cells, map, flags, s, z);
		double alpha = sigma_ps / dyn_dot(nf, z, s);
		muladd(nf, p, p, s, alpha);
		muladd(nf, r, r, z, -alpha);
		double res = r[0]; /* std::max_element: signed maximum, :54 */
		for (size_t k = 1; k < nf; ++k) {
			if (res < r[k]) {
				res = r[k];
			}
		}
		*residual = res;
		if (res < tolerance) {
			++i;
			break;
		}
		mic0_apply(P, a_scale, nf, fluid_cells, map, flags, precon, r, q, z);
		double sigma_new = dyn_dot(nf, z, r);
		double beta = sigma_new / sigma_ps;
		muladd(nf, s, z, s, beta);
		sigma_ps = sigma_new;
	}
	free(precon);
	free(r);
	free(z);
	free(q);
	free(s);
	return i;
}

/* ------------------------------------------------------------------------------------------------ S9 */
void fo_apply_pressure(const fo_params *P, double dt, size_t nf, const uint64_t *fluid_cells,
	const uint64_t *map, const double *p, double *gvel, const uint8_t *type) {
	double coeff = dt / (P->rho * P->h); /* src/pressure_solver.cpp:74 */
	for (size_t i = 0; i < nf; ++i) {
		uint64_t c[3];
		from_raw(P, fluid_cells[i], c);
		double cur = p[i];
		double *v = gvel + 3 * fluid_cells[i];
		for (int d = 0; d < 3; ++d) { /* +faces, :81-124 */
			uint64_t q[3] = { c[0], c[1], c[2] };
			++q[d];
			int in;
			uint8_t t = type_at(P, type, q[0], q[1], q[2], &in);
			if (t != FO_SOLID) {
				double otherp = 0.0;
				if (t == FO_FLUID) {
					otherp = p[map[RAW(P, q[0], q[1], q[2])]];
				}
				v[d] -= coeff * (otherp - cur);
			} else {
				v[d] = 0.0;
			}
		}
		for (int d = 0; d < 3; ++d) { /* -faces owned by non-fluid neighbours, :126-148 */
			uint64_t q[3] = { c[0], c[1], c[2] };
			--q[d];
			int in;
			uint8_t t = type_at(P, type, q[0], q[1], q[2], &in);
			if (in) {
				double *nv = gvel + 3 * RAW(P, q[0], q[1], q[2]);
				if (t == FO_AIR) {
					nv[d] -= coeff * cur;
				} else if (t == FO_SOLID) {
					nv[d] = 0.0;
				}
			}
		}
	}
}

/* ------------------------------------------------------------------------------------------------ E1 */
void fo_extrapolate(const fo_params *P, size_t nf, const uint64_t *fluid_cells, double *gvel, const uint8_t *type) {
	size_t nc = (size_t)(P->nx * P->ny * P->nz);
	const uint64_t size[3] = { P->nx, P->ny, P->nz };
	uint8_t *valid = (uint8_t*)calloc(nc ? nc : 1, 1);
	uint64_t *fresh = (uint64_t*)malloc((nc ? nc : 1) * sizeof(uint64_t));
	size_t nfresh = 0;
	for (size_t i = 0; i < nf; ++i) { /* src/simulation.cpp:686-689 */
		valid[fluid_cells[i]] = 1;
	}
	for (int it = 0; it < P->extrap_iters; ++it) {
		for (size_t k = 0; k < nfresh; ++k) {
			valid[fresh[k]] = 1;
		}
		nfresh = 0;
		for (uint64_t raw = 0; raw < nc; ++raw) { /* valid.for_each: raw order, :701-752 */
			if (valid[raw]) {
				continue;
			}
			uint64_t p[3];
			from_raw(P, raw, p);
			size_t nvalid = 0;
			double nv[3] = { 0.0, 0.0, 0.0 };
			uint8_t type_pos[3] = { FO_SOLID, FO_SOLID, FO_SOLID };
			for (int d = 0; d < 3; ++d) {
				if (p[d] > 0) {
					uint64_t q[3] = { p[0], p[1], p[2] };
					--q[d];
					uint64_t nb = RAW(P, q[0], q[1], q[2]);
					if (valid[nb]) {
						nv[0] += gvel[3 * nb];
						nv[1] += gvel[3 * nb + 1];
						nv[2] += gvel[3 * nb + 2];
						++nvalid;
					}
				}
				if (p[d] + 1 < size[d]) {
					uint64_t q[3] = { p[0], p[1], p[2] };
					++q[d];
					uint64_t nb = RAW(P, q[0], q[1], q[2]);
					if (valid[nb]) {
						nv[0] += gvel[3 * nb];
						nv[1] += gvel[3 * nb + 1];
						nv[2] += gvel[3 * nb + 2];
						type_pos[d] = type[nb];
						++nvalid;
					}
				}
			}
			if (nvalid > 0) {
				for (int d = 0; d < 3; ++d) {
					if (type[raw] == type_pos[d]) {
						gvel[3 * raw + d] = nv[d] / (double)nvalid;
					}
				}
				fresh[nfresh++] = raw;
			}
		}
	}
	free(valid);
	free(fresh);
}

/* ------------------------------------------------------------------------------------------- G1 - G4 */
static inline double lerp1(double a, double b, double t) { /* include/fluid/misc.h:20-22 */
	return a * (1.0 - t) + b * t;
}
static inline double bilerp1(double v00, double v01, double v10, double v11, double t1, double t2) { /* :24-28 */
	return lerp1(lerp1(v00, v01, t2), lerp1(v10, v11, t2), t1);
}
static inline double trilerp1(const double *v, double t1, double t2, double t3) { /* :30-36; v = v000..v111 */
	return lerp1(bilerp1(v[0], v[1], v[2], v[3], t2, t3), bilerp1(v[4], v[5], v[6], v[7], t2, t3), t1);
}
static inline void grad_kernel(const fo_params *P, double px, double py, double pz, double *out) { /* :215-224 */
	double sx = px > 0.0 ? -1.0 : 1.0, sy = py > 0.0 ? -1.0 : 1.0, sz = pz > 0.0 ? -1.0 : 1.0;
	double nx = 1.0 - fabs(px), ny = 1.0 - fabs(py), nz = 1.0 - fabs(pz);
	out[0] = sx * ny * nz / P->h;
	out[1] = nx * sy * nz / P->h;
	out[2] = nx * ny * sz / P->h;
}
static void c_vector(const fo_params *P, const double *v, double tx, double ty, double tz, double *out) { /* :507-521 */
	double acc[3] = { 0.0, 0.0, 0.0 }, g[3];
	for (int k = 0; k < 8; ++k) {
		grad_kernel(P, (k & 1) ? tx - 1.0 : tx, (k & 2) ? ty - 1.0 : ty, (k & 4) ? tz - 1.0 : tz, g);
		double t0 = g[0] * v[k], t1 = g[1] * v[k], t2 = g[2] * v[k];
		if (k == 0) {
			acc[0] = t0;
			acc[1] = t1;
			acc[2] = t2;
		} else {
			acc[0] += t0;
			acc[1] += t1;
			acc[2] += t2;
		}
	}
	out[0] = acc[0];
	out[1] = acc[1];
	out[2] = acc[2];
}
/* mac_grid::get_face_samples (src/mac_grid.cpp:40-112): samples[k][0..7] = component k of v000..v111. */
static void face_samples(const fo_params *P, const double *gvel, const uint64_t *gi, const double *t,
	double samples[3][8], double *tmid) {
	const uint64_t size[3] = { P->nx, P->ny, P->nz };
	uint64_t ci[3][3];
	int clamped[3][3];
	for (int a = 0; a < 3; ++a) {
		for (int d = 0; d < 3; ++d) { /* _clamp(val, 1, max) then -1, :42-50,57-64 */
			uint64_t val = gi[a] + (uint64_t)d;
			if (val < 1) {
				ci[a][d] = 1;
				clamped[a][d] = 1;
			} else if (val >= size[a]) {
				ci[a][d] = size[a];
				clamped[a][d] = 1;
			} else {
				ci[a][d] = val;
				clamped[a][d] = 0;
			}
			--ci[a][d];
		}
	}
	int dsel[3] = { 1, 1, 1 };
	for (int a = 0; a < 3; ++a) { /* :82-95 */
		tmid[a] = t[a] - 0.5;
		if (tmid[a] < 0.0) {
			dsel[a] = 0;
			tmid[a] += 1.0;
		}
	}
#define VEL(k, dx, dy, dz) ((clamped[k][(k) == 0 ? (dx) : (k) == 1 ? (dy) : (dz)]) ? 0.0 : \
	gvel[3 * RAW(P, ci[0][dx], ci[1][dy], ci[2][dz]) + (k)])
	for (int k = 0; k < 8; ++k) { /* :105-112 */
		int bx = k & 1, by = (k >> 1) & 1, bz = (k >> 2) & 1;
		samples[0][k] = VEL(0, bx, dsel[1] + by, dsel[2] + bz);
		samples[1][k] = VEL(1, dsel[0] + bx, by, dsel[2] + bz);
		samples[2][k] = VEL(2, dsel[0] + bx, dsel[1] + by, bz);
	}
#undef VEL
}
static void sample_velocity(const double s[3][8], const double *t, const double *tmid, double *out) {
	out[0] = trilerp1(s[0], tmid[2], tmid[1], t[0]); /* src/simulation.cpp:451-459 */
	out[1] = trilerp1(s[1], tmid[2], t[1], tmid[0]);
	out[2] = trilerp1(s[2], t[2], tmid[1], tmid[0]);
}

void fo_g2p(const fo_params *P, size_t n, const double *pos, double *vel, double *c, const double *gvel,
	const double *old_gvel) {
	for (size_t i = 0; i < n; ++i) {
		uint64_t gi[3];
		double t[3], tmid[3], s[3][8], vnew[3];
		for (int d = 0; d < 3; ++d) { /* compute_cell_index_and_position, src/simulation.cpp:17-23 */
			double f = (pos[3 * i + d] - P->off[d]) / P->h;
			gi[d] = trunc_index(f);
			t[d] = f - (double)gi[d];
		}
		face_samples(P, gvel, gi, t, s, tmid);
		sample_velocity(s, t, tmid, vnew);
		if (P->method == FO_FLIP) { /* :463-505 */
			double so[3][8], vold[3], tm2[3];
			face_samples(P, old_gvel, gi, t, so, tm2);
			sample_velocity(so, t, tmid, vold);
			for (int d = 0; d < 3; ++d) {
				vel[3 * i + d] = vnew[d] + (vel[3 * i + d] - vold[d]) * P->blend;
			}
		} else {
			vel[3 * i] = vnew[0];
			vel[3 * i + 1] = vnew[1];
			vel[3 * i + 2] = vnew[2];
			if (P->method == FO_APIC) { /* :536-544 */
				c_vector(P, s[0], t[0], tmid[1], tmid[2], c + 9 * i);
				c_vector(P, s[1], tmid[0], t[1], tmid[2], c + 9 * i + 3);
				c_vector(P, s[2], tmid[0], tmid[1], t[2], c + 9 * i + 6);
			}
		}
	}
}

/* ------------------------------------------------------------------------------------------- A1 - A4 */
void fo_advect(const fo_params *P, double dt, size_t n, double *pos, const double *vel) {
	const double size[3] = { (double)P->nx, (double)P->ny, (double)P->nz };
	double lo[3], hi[3];
	for (int d = 0; d < 3; ++d) { /* src/simulation.cpp:240-243 */
		lo[d] = P->off[d] + P->skin;
		hi[d] = P->h * size[d] + P->off[d] - P->skin;
	}
	for (size_t i = 0; i < 3 * n; ++i) {
		int d = (int)(i % 3);
		pos[i] = dclamp(pos[i] + vel[i] * dt, lo[d], hi[d]); /* :245-247 */
	}
}

void fo_collide(const fo_params *P, size_t n, double *pos, const double *old_pos, const uint8_t *type) {
	const uint64_t size[3] = { P->nx, P->ny, P->nz };
	const double h = P->h, skin = P->skin;
#pragma omp parallel for schedule(static)
	for (long long ii = 0; ii < (long long)n; ++ii) {
		size_t i = (size_t)ii;
		double from[3] = { old_pos[3 * i], old_pos[3 * i + 1], old_pos[3 * i + 2] };
		double to[3] = { pos[3 * i], pos[3 * i + 1], pos[3 * i + 2] };
		for (int j = 0; j < 3; ++j) { /* src/simulation.cpp:618-651 */
			int into_wall = 0;
			/* grid::march_cells, include/fluid/data_structures/grid.h:140-209 */
			double gf[3], gt[3], diff[3], inv[3], normal[3], t[3];
			int cur[3], tc[3], adv[3];
			for (int d = 0; d < 3; ++d) {
				gf[d] = (from[d] - P->off[d]) / h;
				gt[d] = (to[d] - P->off[d]) / h;
				cur[d] = (int)floor(gf[d]);
				tc[d] = (int)floor(gt[d]);
				diff[d] = gt[d] - gf[d];
				int face;
				if (diff[d] > 0.0) {
					adv[d] = 1;
					face = 1;
				} else {
					adv[d] = -1;
					face = 0;
				}
				inv[d] = 1.0 / fabs(diff[d]);
				normal[d] = -(double)adv[d];
				t[d] = fabs((double)(cur[d] + face) - gf[d]) * inv[d];
			}
			while (cur[0] != tc[0] || cur[1] != tc[1] || cur[2] != tc[2]) {
				int mc = 0;
				double mint = 2.0;
				for (int d = 0; d < 3; ++d) {
					if (t[d] < mint) {
						mint = t[d];
						mc = d;
					}
				}
				if (!(mint <= 1.0)) {
					break;
				}
				cur[mc] += adv[mc];
				int free_cell = 0; /* callback, src/simulation.cpp:621-645 */
				if (cur[0] >= 0 && cur[1] >= 0 && cur[2] >= 0 && (uint64_t)cur[0] < size[0] &&
					(uint64_t)cur[1] < size[1] && (uint64_t)cur[2] < size[2]) {
					free_cell = type[RAW(P, cur[0], cur[1], cur[2])] != FO_SOLID;
				}
				if (!free_cell) {
					double nrm[3] = { 0.0, 0.0, 0.0 }, offv[3] = { to[0] - from[0], to[1] - from[1], to[2] - from[2] };
					nrm[mc] = normal[mc];
					double tt = t[mc] + skin / dot3(offv, nrm);
					tt = dmax(tt, 0.0);
					for (int d = 0; d < 3; ++d) {
						from[d] = tt * to[d] + (1.0 - tt) * from[d];
					}
					to[mc] = from[mc];
					into_wall = 1;
					break;
				}
				t[mc] += inv[mc];
			}
			if (!into_wall) {
				break;
			}
		}
		/* skin push-out, :654-681 (cell index / in-cell position are computed once, before the per-axis pushes) */
		double gp[3], cp[3];
		uint64_t ci[3];
		for (int d = 0; d < 3; ++d) {
			gp[d] = to[d] - P->off[d];
			ci[d] = trunc_index(gp[d] / h);
			cp[d] = gp[d] - (double)ci[d] * h;
		}
		double skin_max = h - skin;
		for (int d = 0; d < 3; ++d) {
			uint64_t q[3] = { ci[0], ci[1], ci[2] };
			if (cp[d] < skin) {
				--q[d];
				if (ci[d] == 0 || type[RAW(P, q[0], q[1], q[2])] == FO_SOLID) {
					to[d] += skin - cp[d];
				}
				++q[d];
			}
			if (cp[d] > skin_max) {
				++q[d];
				if (ci[d] + 1 >= size[d] || type[RAW(P, q[0], q[1], q[2])] == FO_SOLID) {
					to[d] += skin_max - cp[d];
				}
			}
		}
		pos[3 * i] = to[0];
		pos[3 * i + 1] = to[1];
		pos[3 * i + 2] = to[2];
	}
}

static inline uint64_t mix64(uint64_t x) { /* splitmix64 finaliser */
	x ^= x >> 30;
	x *= 0xbf58476d1ce4e5b9ull;
	x ^= x >> 27;
	x *= 0x94d049bb133111ebull;
	x ^= x >> 31;
	return x;
}
void fo_degenerate_kick(const double *p, const double *o, double *out3) {
	uint64_t a[6];
	memcpy(a, p, 24);
	memcpy(a + 3, o, 24);
	uint64_t s = 0x9e3779b97f4a7c15ull;
	for (int k = 0; k < 6; ++k) {
		s = mix64(s ^ a[k]);
	}
	for (int d = 0; d < 3; ++d) {
		s = mix64(s + 0x9e3779b97f4a7c15ull);
		out3[d] = (double)(s >> 11) * (2.0 / 9007199254740992.0) - 1.0; /* uniform in [-1, 1) */
	}
}

void fo_correct(const fo_params *P, double dt, size_t n, double *pos, const uint64_t *begin, const uint64_t *count) {
	const uint64_t size[3] = { P->nx, P->ny, P->nz };
	double re = P->h / sqrt(2.0); /* src/simulation.cpp:566 */
	double *np = (double*)malloc((n ? n : 1) * 3 * sizeof(double));
#pragma omp parallel for schedule(static)
	for (long long ii = 0; ii < (long long)n; ++ii) {
		size_t i = (size_t)ii;
		const double *p = pos + 3 * i;
		uint64_t ci[3];
		for (int d = 0; d < 3; ++d) { /* compute_cell_index, :13-15 */
			ci[d] = trunc_index((p[d] - P->off[d]) / P->h);
		}
		double spring[3] = { 0.0, 0.0, 0.0 };
		uint64_t lo[3], hi[3];
		for (int d = 0; d < 3; ++d) {
			lo[d] = ci[d] < 1 ? 0 : ci[d] - 1;
			hi[d] = ci[d] + 2 < size[d] ? ci[d] + 2 : size[d];
		}
		for (uint64_t cz = lo[2]; cz < hi[2]; ++cz) {
			for (uint64_t cy = lo[1]; cy < hi[1]; ++cy) {
				for (uint64_t cx = lo[0]; cx < hi[0]; ++cx) {
					uint64_t cell = RAW(P, cx, cy, cz);
					for (uint64_t q = begin[cell], k = 0; k < count[cell]; ++q, ++k) {
						if (q == i) {
							continue;
						}
						const double *o = pos + 3 * q;
						double off[3] = { p[0] - o[0], p[1] - o[1], p[2] - o[2] };
						double sq = dot3(off, off);
						if (sq < 1e-12) { /* :584-587 */
							double kick[3];
							fo_degenerate_kick(p, o, kick);
							spring[0] += kick[0];
							spring[1] += kick[1];
							spring[2] += kick[2];
						} else { /* :589-594 */
							double kl = 1.0 - sq / (re * re), kern = 0.0;
							if (kl > 0.0) {
								kern = kl * kl * kl;
							}
							double s = kern / sqrt(sq);
							spring[0] += s * off[0];
							spring[1] += s * off[1];
							spring[2] += s * off[2];
						}
					}
				}
			}
		}
		double f = dt * P->stiffness * re; /* :599 */
		for (int d = 0; d < 3; ++d) {
			np[3 * i + d] = p[d] + spring[d] * f;
		}
	}
	for (size_t i = 0; i < n; ++i) { /* :604-608 */
		for (int d = 0; d < 3; ++d) {
			double gmax = P->off[d] + (double)size[d] * P->h;
			pos[3 * i + d] = dclamp(np[3 * i + d], P->off[d], gmax);
		}
	}
	free(np);
}

double fo_cfl(const fo_params *P, size_t n, const double *vel) {
	double maxlen = 0.0;
	for (size_t i = 0; i < n; ++i) {
		maxlen = dmax(maxlen, dot3(vel + 3 * i, vel + 3 * i));
	}
	return P->h / sqrt(maxlen);
}

/* -------------------------------------------------------------------------------------- whole step */
static double now_s(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static void permute_rows(size_t n, size_t w, double *a, const uint64_t *perm, double *tmp) {
	for (size_t i = 0; i < n; ++i) {
		memcpy(tmp + w * i, a + w * perm[i], w * sizeof(double));
	}
	memcpy(a, tmp, n * w * sizeof(double));
}

size_t fo_time_step(const fo_params *P, double dt, size_t n, double *pos, double *vel, double *c, double *old_pos,
	double *gvel, uint8_t *type, double *old_gvel, double tolerance, size_t max_iterations, double *residual,
	double *ph) {
	size_t nc = (size_t)(P->nx * P->ny * P->nz), na = n ? n : 1;
	uint64_t *key = (uint64_t*)malloc(na * 8), *perm = (uint64_t*)malloc(na * 8);
	uint64_t *begin = (uint64_t*)malloc(nc * 8), *count = (uint64_t*)malloc(nc * 8);
	uint64_t *fluid = (uint64_t*)malloc((nc < na ? nc : na) * 8), *map = (uint64_t*)malloc(nc * 8);
	double *tmp = (double*)malloc(na * 9 * 8);
	double t0, acc[8] = { 0 };
	/* the first update_and_hash_particles (src/simulation.cpp:49) only feeds source coercion: skipped */
	t0 = now_s();
	fo_advect(P, dt, n, pos, vel);
	fo_collide(P, n, pos, old_pos, type);
	memcpy(old_pos, pos, n * 24);
	acc[1] += now_s() - t0;
	t0 = now_s();
	fo_cell_keys(P, n, pos, key);
	size_t nf = fo_hash(P, n, key, perm, begin, count, fluid);
	permute_rows(n, 3, pos, perm, tmp);
	permute_rows(n, 3, vel, perm, tmp);
	permute_rows(n, 9, c, perm, tmp);
	memcpy(old_pos, pos, n * 24);
	acc[0] += now_s() - t0;
	t0 = now_s();
	fo_p2g(P, n, pos, vel, c, begin, count, gvel, type, old_gvel);
	fo_gravity(P, dt, gvel);
	acc[2] += now_s() - t0;
	t0 = now_s();
	uint8_t *flags = (uint8_t*)malloc(nf ? nf : 1);
	double *b = (double*)malloc((nf ? nf : 1) * 8), *p = (double*)malloc((nf ? nf : 1) * 8);
	fo_solver_setup(P, gvel, type, nf, fluid, map, flags, b);
	size_t iters = fo_solve(P, dt, nf, fluid, map, flags, b, 0.97, 0.25, tolerance, max_iterations, p, residual);
	acc[3] += now_s() - t0;
	t0 = now_s();
	fo_apply_pressure(P, dt, nf, fluid, map, p, gvel, type);
	acc[4] += now_s() - t0;
	t0 = now_s();
	fo_correct(P, dt, n, pos, begin, count);
	fo_collide(P, n, pos, old_pos, type);
	memcpy(old_pos, pos, n * 24);
	acc[5] += now_s() - t0;
	t0 = now_s();
	fo_extrapolate(P, nf, fluid, gvel, type);
	acc[6] += now_s() - t0;
	t0 = now_s();
	fo_g2p(P, n, pos, vel, c, gvel, old_gvel);
	acc[7] += now_s() - t0;
	if (ph) {
		for (int k = 0; k < 8; ++k) {
			ph[k] += acc[k];
		}
	}
	free(key); free(perm); free(begin); free(count); free(fluid); free(map); free(tmp);
	free(flags); free(b); free(p);
	return iters;
}
