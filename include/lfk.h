/* lfk -- "libfluid kernels": the C ABI between a libfluid-compatible host (the fluid::simulation /
 * fluid::mac_grid / fluid::pressure_solver shim in include/fluid/, or any FFI) and the B200-native CUDA
 * implementation of libfluid's per-step hot path.
 *
 * The reference (lukedan/libfluid) has no FFI layer: its boundary is the public C++ class API.  Every entry point
 * below therefore names the reference member function it replaces (file:line in the reference tree).
 *
 * Conventions
 *   - every function returns 0 on success and a negative code on failure (-(cudaError_t) for CUDA errors,
 *     -(1000 + ncclResult_t) for NCCL errors, LFK_E_* otherwise); nothing throws, nothing aborts.  The message of
 *     the last failure is kept per context (lfk_last_error).
 *   - host pointers are borrowed for the duration of the call only.  Device state (SoA particles, SoA face
 *     velocities, cell types, solver vectors) is owned by the opaque lfk_ctx.
 *   - one context <-> one GPU <-> one host thread at a time.  Multi-GPU = one context per rank (z-slab
 *     decomposition); the ranks are tied together with an NCCL unique id (lfk_nccl_unique_id).
 *   - there is NO CPU fallback: without a CUDA device lfk_create fails with LFK_E_NO_DEVICE.
 *   - host layouts are exactly the reference's: particles are the 152-byte records of
 *     fluid::simulation::particle (include/fluid/simulation.h:24-34: position, velocity, cx, cy, cz,
 *     old_position as 3 doubles each, then size_t raw_cell_index), cells are the 32-byte records of
 *     fluid::mac_grid::cell (include/fluid/mac_grid.h:15-27: vec3d velocities_posface, 1-byte type, padding),
 *     cells in raw order x + nx*(y + ny*z) (include/fluid/data_structures/grid.h:212-222).
 */
#ifndef LFK_H
#define LFK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LFK_ABI_VERSION 1

typedef struct lfk_ctx lfk_ctx;

enum { LFK_CELL_AIR = 1, LFK_CELL_FLUID = 2, LFK_CELL_SOLID = 4 };     /* mac_grid::cell::type, mac_grid.h:17-21 */
enum { LFK_METHOD_PIC = 0, LFK_METHOD_FLIP = 1, LFK_METHOD_APIC = 2 }; /* simulation::method, simulation.h:44-48 */
enum { LFK_PRECOND_JACOBI = 0, LFK_PRECOND_MULTIGRID = 1 };

enum {
	LFK_OK = 0,
	LFK_E_INVALID = -2000,   /* bad argument */
	LFK_E_NO_DEVICE = -2001, /* no CUDA device / CUDA runtime unusable */
	LFK_E_STATE = -2002,     /* call made in the wrong state (e.g. apply_pressure before a solve) */
	LFK_E_CAPACITY = -2003,  /* caller buffer too small */
	LFK_E_NCCL = -2004       /* built without NCCL but nranks > 1 */
};

/* The public data members of fluid::simulation (simulation.h:177-190) and of fluid::pressure_solver
 * (pressure_solver.h:38-42) that steer the hot path, as one POD. */
typedef struct lfk_params {
	double grid_offset[3];      /* simulation::grid_offset */
	double cell_size;           /* simulation::cell_size (must be set; the reference defaults to NaN) */
	double density;             /* simulation::density */
	double gravity[3];          /* simulation::gravity */
	double boundary_skin_width; /* simulation::boundary_skin_width */
	double correction_stiffness;/* simulation::correction_stiffness */
	double blending_factor;     /* simulation::blending_factor (1.0 = pure FLIP) */
	double cfl_number;          /* simulation::cfl_number */
	double tolerance;           /* pressure_solver::tolerance (1e-6) */
	int32_t method;             /* simulation::simulation_method, LFK_METHOD_* */
	int32_t extrapolation_iterations; /* simulation::velocity_extrapolation_iterations */
	int32_t max_iterations;     /* pressure_solver::max_iterations (200) */
	int32_t preconditioner;     /* LFK_PRECOND_*; replaces the reference's sequential MIC(0) */
} lfk_params;

/* Counters and device timings (CUDA events on the context's stream) of the most recent calls. */
typedef struct lfk_stats {
	uint64_t kernel_launches;   /* kernels launched by this context since creation / last reset */
	uint64_t pcg_iterations;    /* of the last lfk_pressure_solve */
	double pcg_residual;        /* max |r_i| of the last solve */
	double phase_ms[16];        /* accumulated per LFK_PHASE_* since the last reset (only when timing is on) */
	uint64_t num_particles;
	uint64_t num_fluid_cells;
	uint64_t exchanged_particles; /* multi-GPU: particles sent to the z neighbours by the last sort */
} lfk_stats;
enum {
	LFK_PHASE_ADVECT_COLLIDE = 0, LFK_PHASE_SORT = 1, LFK_PHASE_P2G = 2, LFK_PHASE_SOLVE_SETUP = 3,
	LFK_PHASE_PCG = 4, LFK_PHASE_APPLY_PRESSURE = 5, LFK_PHASE_CORRECT_COLLIDE = 6, LFK_PHASE_EXTRAPOLATE = 7,
	LFK_PHASE_G2P = 8, LFK_PHASE_CFL = 9, LFK_PHASE_TRANSFER = 10, LFK_PHASE_EXCHANGE = 11, LFK_PHASE_COUNT = 12
};

/* ---- lifetime ------------------------------------------------------------------------------------------- */
int lfk_abi_version(void);
/* Fills 128 bytes with an NCCL unique id (rank 0 calls it; the host distributes it to the other ranks). */
int lfk_nccl_unique_id(void *out128);
/* simulation::resize (src/simulation.cpp:26-29).  nx,ny,nz: global grid.  device: CUDA ordinal.  stream: a
 * cudaStream_t to run on, or NULL for a private stream.  nranks/rank/nccl_id: z-slab decomposition over
 * nranks GPUs (nranks = 1: nccl_id may be NULL). */
int lfk_create(lfk_ctx **out, uint64_t nx, uint64_t ny, uint64_t nz, int device, void *stream,
	int nranks, int rank, const void *nccl_id128);
int lfk_destroy(lfk_ctx *ctx);
const char *lfk_last_error(const lfk_ctx *ctx); /* ctx may be NULL: error of the last failed lfk_create */
int lfk_set_params(lfk_ctx *ctx, const lfk_params *p);
int lfk_get_params(const lfk_ctx *ctx, lfk_params *p);
int lfk_sync(lfk_ctx *ctx);
/* z range [z_begin, z_end) of the cells this rank owns. */
int lfk_slab(const lfk_ctx *ctx, uint64_t *z_begin, uint64_t *z_end);

/* ---- state transfer (host mirrors of simulation::particles() / simulation::grid()) ----------------------- */
/* simulation::particles() = ... (simulation.h:142-148).  Multi-GPU: every rank passes the particles it wants to
 * contribute; ownership is settled by the next lfk_hash / lfk_time_step. */
int lfk_upload_particles(lfk_ctx *ctx, const void *aos152, uint64_t n);
int lfk_num_particles(lfk_ctx *ctx, uint64_t *n);
int lfk_download_particles(lfk_ctx *ctx, void *aos152, uint64_t capacity, uint64_t *n);
/* positions only, 24 B per particle (what the mesher / renderer / Maya node consume every frame) */
int lfk_download_positions(lfk_ctx *ctx, double *xyz, uint64_t capacity, uint64_t *n);
/* The same, asynchronously: returns as soon as the copy is queued on the context's transfer stream; the next stage
 * calls may be issued at once and overlap the copy.  xyz must stay valid -- and should be pinned (lfk_host_alloc) -- until
 * lfk_wait_transfers returns.  This is the per-frame read of testbed/main.cpp:52 and grid_node.cpp:358-366. */
int lfk_download_positions_async(lfk_ctx *ctx, double *xyz, uint64_t capacity, uint64_t *n);
int lfk_wait_transfers(lfk_ctx *ctx);
/* page-locked host memory for hosts that do not link the CUDA runtime themselves */
int lfk_host_alloc(void **out, uint64_t bytes);
int lfk_host_free(void *p);
/* Binary checkpoint of the rank's state (particles as SoA fields, the slab's cells, the solver's warm-start state);
 * multi-GPU: one file per rank, `path` + ".rank<r>".  Loading it into a context of the same grid and rank layout
 * continues the run bit for bit.  Replaces the text point cloud of include/fluid/data_structures/point_cloud.h:14-37
 * for restarts (the host mirror still offers save_to_naive / load_from_naive for interchange). */
int lfk_checkpoint_save(lfk_ctx *ctx, const char *path);
int lfk_checkpoint_load(lfk_ctx *ctx, const char *path);
/* simulation::grid().grid() (mac_grid.h:62-69): the whole nx*ny*nz grid; each rank keeps its slab. */
int lfk_upload_cells(lfk_ctx *ctx, const void *aos32);
/* writes the cells this rank owns into the whole-grid array (other entries untouched) */
int lfk_download_cells(lfk_ctx *ctx, void *aos32);
/* Multi-GPU hosts that mirror only their own slab: the same two transfers on slab-sized buffers.  Upload: the buffer
 * holds the layers [max(z_begin - 1, 0), min(z_end + 1, nz)) (the slab plus its in-domain ghost layers); download: the
 * owned layers [z_begin, z_end) (lfk_slab). */
int lfk_upload_cells_slab(lfk_ctx *ctx, const void *aos32_slab);
int lfk_download_cells_slab(lfk_ctx *ctx, void *aos32_own);
/* simulation::_old_grid (simulation.h:201-202), FLIP only */
int lfk_upload_old_cells(lfk_ctx *ctx, const void *aos32);
int lfk_download_old_cells(lfk_ctx *ctx, void *aos32);
/* simulation::_space_hash / _fluid_cells (simulation.h:203-209): begin/count per cell of this rank's slab in
 * whole-grid raw indexing (begin relative to this rank's particle array). */
int lfk_download_table(lfk_ctx *ctx, uint64_t *begin, uint64_t *count);
int lfk_num_fluid_cells(lfk_ctx *ctx, uint64_t *nf);
int lfk_download_fluid_cells(lfk_ctx *ctx, uint64_t *raw, uint64_t capacity);

/* ---- stages of simulation::time_step (src/simulation.cpp:43-125), one call each -------------------------- */
/* update_and_hash_particles (src/simulation.cpp:251-291): keys, stable cell sort, {begin,count} table */
int lfk_hash(lfk_ctx *ctx);
/* _advect_particles without sources (src/simulation.cpp:240-248) */
int lfk_advect(lfk_ctx *ctx, double dt);
/* _detect_collisions + old_position = position (src/simulation.cpp:612-683, 56-59) */
int lfk_collide(lfk_ctx *ctx);
/* _transfer_to_grid (src/simulation.cpp:293-412, 428-445); needs a valid table (lfk_hash) */
int lfk_p2g(lfk_ctx *ctx);
/* the gravity loop (src/simulation.cpp:72-78) */
int lfk_gravity(lfk_ctx *ctx, double dt);
/* pressure_solver::solve (src/pressure_solver.cpp:19-71): converges to max|r| < tolerance in the reference's
 * scaling of A and b; iteration counts differ from the reference's MIC(0). */
int lfk_pressure_solve(lfk_ctx *ctx, double dt, double *residual, uint64_t *iterations);
/* the solver's b vector and matrix flags in fluid-cell order (src/pressure_solver.cpp:157-242); flags =
 * nonsolid_neighbors | fluid_xpos<<3 | fluid_ypos<<4 | fluid_zpos<<5.  Either pointer may be NULL. */
int lfk_download_rhs(lfk_ctx *ctx, double dt, double *b, uint8_t *flags, uint64_t capacity);
/* the pressure vector in fluid-cell order (ascending raw index), as solve() returns it */
int lfk_download_pressure(lfk_ctx *ctx, double *p, uint64_t capacity);
int lfk_upload_pressure(lfk_ctx *ctx, const double *p, uint64_t n);
/* out = A v with the reference's scaling (_apply_a, src/pressure_solver.cpp:334-362); v, out in fluid-cell order */
int lfk_apply_a(lfk_ctx *ctx, double dt, const double *v, double *out, uint64_t n);
/* pressure_solver::apply_pressure (src/pressure_solver.cpp:73-148) */
int lfk_apply_pressure(lfk_ctx *ctx, double dt);
/* _correct_positions (src/simulation.cpp:562-610); needs the table of the last lfk_hash */
int lfk_correct(lfk_ctx *ctx, double dt);
/* _extrapolate_velocities (src/simulation.cpp:685-754) */
int lfk_extrapolate(lfk_ctx *ctx);
/* _transfer_from_grid (src/simulation.cpp:447-560, src/mac_grid.cpp:40-112) */
int lfk_g2p(lfk_ctx *ctx);
/* simulation::cfl (src/simulation.cpp:199-205).  After a G2P (staged or inside lfk_time_step) the maximum |v|^2 comes
 * from that kernel; anything that rewrites particles afterwards (uploads, seeding, sources, checkpoint load) makes the
 * next call reduce over the particles again.  The value is the same either way (a maximum has no order). */
int lfk_cfl(lfk_ctx *ctx, double *value);

/* ---- fluid sources (simulation::sources, include/fluid/data_structures/source.h:12-22) on the device --------- */
typedef struct lfk_source {
	const uint64_t *cells;      /* source::cells as x, y, z triples (whole-grid cell indices) */
	uint64_t num_cells;
	double velocity[3];         /* source::velocity */
	uint32_t target_density_cubic_root; /* source::target_density_cubic_root */
	int32_t active;             /* source::active */
	int32_t coerce_velocity;    /* source::coerce_velocity */
} lfk_source;
/* Replaces the context's source list (n = 0 clears it).  While at least one source is active, lfk_time_step runs the
 * reference's source handling on the device: velocity coercion before advection, seed_cell after the first sort of the
 * step, and a second sort when particles were added (src/simulation.cpp:49-64). */
int lfk_set_sources(lfk_ctx *ctx, const lfk_source *sources, uint64_t n);
/* Seed of the counter-based generator that places source particles.  The reference draws from the simulation's pcg32
 * (simulation.h:176); the device draws uniform positions in the cell from a hash of (seed, step, cell, slot): same
 * distribution, same counts and velocities, different positions. */
int lfk_set_rng_seed(lfk_ctx *ctx, uint64_t seed);
/* the coercion half of _advect_particles (src/simulation.cpp:227-238): particles whose cell belongs to an active
 * source with coerce_velocity get the source's velocity and zero APIC rows */
int lfk_coerce_sources(lfk_ctx *ctx);
/* _update_sources + seed_cell (src/simulation.cpp:756-765, 136-151): every cell of every active source is filled up
 * to target_density_cubic_root^3 particles; needs the table of the last lfk_hash, invalidates it when *added > 0 */
int lfk_update_sources(lfk_ctx *ctx, uint64_t *added);

/* ---- neighbours of the path that consume / produce its device-resident state ------------------------------ */
/* mesher (include/fluid/mesher.h:15-46): the public fields and resize() */
typedef struct lfk_mesher {
	double grid_offset[3];   /* mesher::grid_offset */
	double cell_size;        /* mesher::cell_size */
	double particle_extent;  /* mesher::particle_extent */
	uint64_t cell_radius;    /* mesher::cell_radius */
	uint64_t size[3];        /* mesher::resize(size): cells; the surface function has size + 1 samples per axis */
} lfk_mesher;
/* mesher::_sample_surface_function (src/mesher.cpp:333-376), the first half of generate_mesh(particles, r): the
 * implicit surface function on the (size + 1)^3 sample points, x fastest.  xyz: host positions (x, y, z triples), or
 * NULL to sample the context's own particles where they live (no position download).  Points without particles in
 * range read 1.0, points whose particles all weigh zero read NaN, like the reference. */
int lfk_mesher_sample(lfk_ctx *ctx, const lfk_mesher *m, double r, const double *xyz, uint64_t n, double *surface);
/* voxelizer::resize_reposition_grid_constrained + voxelize_mesh_surface + mark_exterior (src/voxelizer.cpp:19-126) for
 * a triangle mesh (positions: x, y, z triples; indices: 3 per triangle).  Returns the voxel grid's offset in cells of
 * the reference grid and its size; the voxels stay on the device. */
int lfk_voxelize_mesh(lfk_ctx *ctx, const double *positions, uint64_t num_vertices, const uint64_t *indices,
	uint64_t num_indices, double cell_size, const double ref_grid_offset[3], int64_t grid_min[3], uint64_t voxel_size[3]);
/* voxelizer::voxels as bytes (0 interior, 1 exterior, 2 surface: voxelizer::cell_type), x fastest */
int lfk_voxels_download(lfk_ctx *ctx, uint8_t *voxels, uint64_t capacity);
/* obstacle::cells (src/data_structures/obstacle.cpp:9-29): the interior voxels that lie inside the simulation grid, as
 * x, y, z triples in the reference's order; cells_xyz may be NULL (count only).  mark_solid != 0 also sets their cell
 * type to solid on the device (what plugins/maya/nodes/grid_node.cpp:330-340 does cell by cell on the host).
 * Where the voxel grid starts above the simulation grid's origin the reference's own loop bounds mix the two grids'
 * coordinates and read out of bounds; this is the intended set. */
int lfk_obstacle_cells(lfk_ctx *ctx, uint64_t *cells_xyz, uint64_t capacity, uint64_t *n, int mark_solid);

/* ---- fused entry points (what the shim calls when no mid-step callback is installed) --------------------- */
/* simulation::time_step(dt) without sources (src/simulation.cpp:43-125), entirely on the device */
int lfk_time_step(lfk_ctx *ctx, double dt);
/* simulation::time_step() (src/simulation.cpp:127-129): dt = min(cfl_number * cfl(), 0.033); returns dt used */
int lfk_time_step_cfl(lfk_ctx *ctx, double *dt_used);
/* simulation::update(dt) (src/simulation.cpp:31-41); returns the number of sub-steps taken */
int lfk_update(lfk_ctx *ctx, double dt, uint64_t *substeps);

/* ---- synthetic scenes for benchmarks (device-side; jittered-subcell seeding with the distribution of
 * simulation::seed_func, simulation.h:80-115, from a counter-based hash RNG -- not bit-compatible with pcg32) - */
int lfk_seed_box_device(lfk_ctx *ctx, const double start[3], const double size[3], const double velocity[3],
	uint32_t density, uint64_t seed, int append);
/* projection-only benchmark: all-fluid box, air top layer, i.i.d. uniform(-1,1) face velocities (config 5) */
int lfk_synthetic_projection_device(lfk_ctx *ctx, uint64_t seed);

/* ---- instrumentation ------------------------------------------------------------------------------------ */
int lfk_set_timing(lfk_ctx *ctx, int enabled);
/* A/B switches between code paths that compute the same result (profiling aid; every default is the production
 * path).  Keys: "p2g" 0 z-marching kernel / 2 plain per-cell gather (the reference's loop literally); "lean_sort" 1 the
 * fused step permutes positions only and P2G reads velocity / c rows through the permutation / 0 full-payload sort;
 * "warm_start" 1 the fused step starts PCG from the previous pressure / 0 from p = 0 like the reference; "graph" 1 the
 * PCG iteration is replayed from a captured CUDA graph / 0 launched kernel by kernel; "red_blocks" n caps the grid of
 * the PCG reduction kernels; "mg_coarse" n symmetric sweeps on the coarsest multigrid level (default 8); "p2g_chunk" n
 * z planes per block of the marching P2G kernel (8 .. 128, default: about 6 waves of one block per SM).
 * Multi-GPU (set them on every rank alike): "p2p" 1 halos and PCG scalars through CUDA-IPC peer memory / 0 NCCL;
 * "mg_agg" 1 coarse multigrid levels agglomerated onto every rank / 0 distributed; "mg_agg_cells" n largest whole-grid
 * level that is agglomerated (default 600000); "ll_kb" n halo layers up to n KB (<= 2048, default 1024) travel as
 * {element, epoch} words (no fence, no flag round trip).  The environment variable LFK_TUNE="key=value,..." applies the same
 * switches to every context of the process.  Unknown key: LFK_E_INVALID. */
int lfk_set_tuning(lfk_ctx *ctx, const char *key, int value);
int lfk_get_stats(lfk_ctx *ctx, lfk_stats *out);
int lfk_reset_stats(lfk_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* LFK_H */
