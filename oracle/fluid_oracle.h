/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C11) of the hot path of lukedan/libfluid.
 *
 * Every function cites the reference file:line it follows.  The restatement is pinned bit-for-bit against the
 * compiled reference (oracle/_ref, see tests/test_oracle_pin.py) and against the golden vectors minted from
 * it (tests/golden/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call it; the
 * product path (libfluid_b200/) never does.
 *
 * Layouts (host, contiguous): per-particle vectors are [n][3] doubles (APIC c is [n][9] = rows cx,cy,cz), grid
 * velocities are [ncells][3] doubles with the raw cell index x + nx*(y + ny*z) of the reference
 * (include/fluid/data_structures/grid.h:24-31,212-222), cell types are one byte (air 1, fluid 2, solid 4;
 * include/fluid/mac_grid.h:17-21).
 */
#ifndef FLUID_ORACLE_H
#define FLUID_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { FO_AIR = 1, FO_FLUID = 2, FO_SOLID = 4 };
enum { FO_PIC = 0, FO_FLIP = 1, FO_APIC = 2 };
#define FO_NOT_FLUID UINT64_MAX

typedef struct fo_params {
	uint64_t nx, ny, nz;
	double h;            /* cell_size */
	double off[3];       /* grid_offset */
	double g[3];         /* gravity */
	double rho;          /* density */
	double skin;         /* boundary_skin_width */
	double stiffness;    /* correction_stiffness */
	double blend;        /* blending_factor */
	int32_t method;      /* FO_PIC / FO_FLIP / FO_APIC */
	int32_t extrap_iters;/* velocity_extrapolation_iterations */
} fo_params;

/* K1  src/simulation.cpp:251-261 */
void fo_cell_keys(const fo_params *P, size_t n, const double *pos, uint64_t *key);
/* K2  src/simulation.cpp:266-291; stable counting sort (the reference's std::sort leaves in-cell order
 * unspecified).  perm[i] = source index of the i-th particle in sorted order.  Returns the number of fluid cells. */
size_t fo_hash(const fo_params *P, size_t n, const uint64_t *key, uint64_t *perm, uint64_t *begin,
	uint64_t *count, uint64_t *fluid_cells);
/* Same table from keys that are already sorted (src/simulation.cpp:273-290). */
size_t fo_cell_ranges(const fo_params *P, size_t n, const uint64_t *sorted_key, uint64_t *begin, uint64_t *count,
	uint64_t *fluid_cells);
/* P1-P3  src/simulation.cpp:293-412,428-445.  old_gvel is written only for FO_FLIP (may be NULL otherwise). */
void fo_p2g(const fo_params *P, size_t n, const double *pos, const double *vel, const double *c,
	const uint64_t *begin, const uint64_t *count, double *gvel, uint8_t *type, double *old_gvel);
/* G0  src/simulation.cpp:72-78 */
void fo_gravity(const fo_params *P, double dt, double *gvel);
/* S1-S3  src/pressure_solver.cpp:150-242.  index_map[ncells] (FO_NOT_FLUID elsewhere), flags[nf] =
 * nonsolid | xpos<<3 | ypos<<4 | zpos<<5, b[nf]. */
void fo_solver_setup(const fo_params *P, const double *gvel, const uint8_t *type, size_t nf,
	const uint64_t *fluid_cells, uint64_t *index_map, uint8_t *flags, double *b);
/* S6  src/pressure_solver.cpp:334-362 */
void fo_apply_a(const fo_params *P, double a_scale, size_t nf, const uint64_t *fluid_cells,
	const uint64_t *index_map, const uint8_t *flags, const double *v, double *out);
/* S4-S8  src/pressure_solver.cpp:19-71,244-332,364-370.  MIC(0)-PCG.  Returns iterations; p[nf]. */
size_t fo_solve(const fo_params *P, double dt, size_t nf, const uint64_t *fluid_cells, const uint64_t *index_map,
	const uint8_t *flags, const double *b, double tau, double sigma, double tolerance, size_t max_iterations,
	double *p, double *residual);
/* S9  src/pressure_solver.cpp:73-148 */
void fo_apply_pressure(const fo_params *P, double dt, size_t nf, const uint64_t *fluid_cells,
	const uint64_t *index_map, const double *p, double *gvel, const uint8_t *type);
/* E1  src/simulation.cpp:685-754 */
void fo_extrapolate(const fo_params *P, size_t nf, const uint64_t *fluid_cells, double *gvel, const uint8_t *type);
/* G1-G4  src/mac_grid.cpp:40-112, src/simulation.cpp:447-560.  old_gvel only for FO_FLIP. */
void fo_g2p(const fo_params *P, size_t n, const double *pos, double *vel, double *c, const double *gvel,
	const double *old_gvel);
/* A1  src/simulation.cpp:240-248 (no sources) */
void fo_advect(const fo_params *P, double dt, size_t n, double *pos, const double *vel);
/* A2  src/simulation.cpp:612-683 + include/fluid/data_structures/grid.h:140-209 */
void fo_collide(const fo_params *P, size_t n, double *pos, const double *old_pos, const uint8_t *type);
/* A3  src/simulation.cpp:562-610.  The reference's r^2<1e-12 branch draws from std::random_device and is not
 * reproducible; here it is a deterministic hash kick (same distribution), see fo_degenerate_kick. */
void fo_correct(const fo_params *P, double dt, size_t n, double *pos, const uint64_t *begin, const uint64_t *count);
void fo_degenerate_kick(const double *p, const double *o, double *out3);
/* A4  src/simulation.cpp:199-205 */
double fo_cfl(const fo_params *P, size_t n, const double *vel);

/* Whole time_step without sources (src/simulation.cpp:43-125), composed from the pieces above; particle arrays are
 * permuted in place by the (stable) sort.  Scratch is allocated internally. Returns PCG iterations. */
size_t fo_time_step(const fo_params *P, double dt, size_t n, double *pos, double *vel, double *c, double *old_pos,
	double *gvel, uint8_t *type, double *old_gvel, double tolerance, size_t max_iterations, double *residual,
	double *phase_seconds /* [8] or NULL: sort, advect+collide, p2g, solve, apply, correct, extrapolate, g2p */);

#ifdef __cplusplus
}
#endif
#endif
