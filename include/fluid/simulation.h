#pragma once
// fluid::simulation -- source-compatible mirror of the reference class (include/fluid/simulation.h:20-281) whose
// per-step hot path runs on a B200 through the lfk C ABI (include/lfk.h).
//
// Same public surface as the reference: particle record, method enum, resize / update / time_step, hashing and
// seeding helpers, cfl(), grid() / particles() accessors, the eight step callbacks, and the public tuning members.
// What changes is ownership: the authoritative particle and grid state lives on the device (SoA); `particles()` and
// `grid()` hand out references to HOST MIRRORS that are synchronised lazily:
//   * a non-const accessor downloads the mirror if the device copy is newer and marks the device copy stale, so
//     whatever the caller writes is uploaded before the next device stage;
//   * when no mid-step callback is installed and no source is active, time_step() is ONE fused device call
//     (lfk_time_step) and nothing crosses PCIe;
//   * installed callbacks fire at the reference's points (src/simulation.cpp:43-125) and may read any state.
// Seeding and sources stay on the host (src/simulation.cpp:136-151, 756-765), with the reference's RNG stream.
#include <functional>
#include <limits>
#include <memory>
#include <random>
#include <vector>

#include "data_structures/grid.h"
#include "data_structures/source.h"
#include "mac_grid.h"
#include "pcg32.h"

struct lfk_ctx;

namespace fluid {
	class pressure_solver;

	class simulation {
		friend class pressure_solver;
	public:
		/// A particle: exactly the 152-byte record of the reference (include/fluid/simulation.h:24-34).
		struct particle {
			vec3d position, velocity, cx, cy, cz;
			vec3d old_position;
			std::size_t raw_cell_index = 0;

			vec3s compute_cell_index(vec3d grid_offset, double cell_size) const {
				return vec3s((position - grid_offset) / cell_size);
			}
			std::pair<vec3s, vec3d> compute_cell_index_and_position(vec3d grid_offset, double cell_size) const {
				vec3d f = (position - grid_offset) / cell_size;
				vec3s i(f);
				return { i, f - vec3d(i) };
			}
		};
		static_assert(sizeof(particle) == 152, "lfk_upload_particles expects the reference's 152-byte records");
		enum class method : unsigned char { pic, flip_blend, apic };
		/// preconditioner of the device pressure solver (replaces the reference's sequential MIC(0))
		enum class preconditioner : unsigned char { jacobi, multigrid };

		constexpr static bool precise_collision_detection = true;
		constexpr static std::size_t default_seeding_density = 2;

		simulation();
		~simulation();
		simulation(const simulation&) = delete;
		simulation &operator=(const simulation&) = delete;

		void resize(vec3s);
		void update(double dt);
		void time_step(double dt);
		void time_step();

		void reset_space_hash();
		void update_and_hash_particles();
		void hash_particles();

		void seed_cell(vec3s cell, vec3d velocity, std::size_t density = default_seeding_density);
		template <typename Func> void seed_func(
			vec3s start, vec3s size, const Func &pred, vec3d velocity = vec3d(),
			std::size_t density = default_seeding_density
		) {
			std::vector<particle> &ps = particles();
			double sub = cell_size / static_cast<double>(density);
			std::uniform_real_distribution<double> dist(0.0, sub);
			vec3s gs = grid().grid().get_size(), end;
			for (std::size_t d = 0; d < 3; ++d) {
				end[d] = start[d] + size[d] < gs[d] ? start[d] + size[d] : gs[d];
			}
			for (std::size_t z = start.z; z < end.z; ++z) {
				for (std::size_t y = start.y; y < end.y; ++y) {
					for (std::size_t x = start.x; x < end.x; ++x) {
						vec3d cell_offset = vec3d(vec3s(x, y, z)) * cell_size;
						std::size_t raw = grid().grid().index_to_raw(vec3s(x, y, z));
						for (std::size_t sx = 0; sx < density; ++sx) {
							for (std::size_t sy = 0; sy < density; ++sy) {
								for (std::size_t sz = 0; sz < density; ++sz) {
									// Three draws per candidate (also for rejected points).  The reference writes
									// vec3d(dist(random), dist(random), dist(random)) -- an unspecified evaluation order
									// that g++ resolves right to left: z is drawn first, x last.  Same order here, so
									// the particle stream is bit-identical to the reference as built by its toolchain.
									double jz = dist(random), jy = dist(random), jx = dist(random);
									vec3d pos = grid_offset + cell_offset + vec3d(vec3s(sx, sy, sz)) * sub + vec3d(jx, jy, jz);
									if (pred(pos)) {
										particle p;
										p.old_position = p.position = pos;
										p.velocity = velocity;
										p.raw_cell_index = raw;
										ps.emplace_back(p);
									}
								}
							}
						}
					}
				}
			}
		}
		void seed_box(vec3d start, vec3d size, vec3d velocity = vec3d(), std::size_t density = default_seeding_density);
		void seed_sphere(vec3d center, double radius, vec3d velocity = vec3d(), std::size_t density = default_seeding_density);

		[[nodiscard]] vec3s world_position_to_cell_index(vec3d) const;
		[[nodiscard]] vec3s world_position_to_cell_index_unclamped(vec3d) const;
		[[nodiscard]] double cfl() const;

		/// Mutable access: the host mirror is refreshed first and the device copy is considered stale afterwards.
		[[nodiscard]] mac_grid &grid();
		[[nodiscard]] const mac_grid &grid() const;
		[[nodiscard]] std::vector<particle> &particles();
		[[nodiscard]] const std::vector<particle> &particles() const;

		// callbacks, in calling order (reference include/fluid/simulation.h:150-175)
		std::function<void(double)> pre_time_step_callback;
		std::function<void(double)> post_advection_callback;
		std::function<void(double)> post_particle_to_grid_transfer_callback;
		std::function<void(double)> post_gravity_callback;
		std::function<void(double, std::vector<double>&, double, std::size_t)> post_pressure_solve_callback;
		std::function<void(double)> post_apply_pressure_callback;
		std::function<void(double)> post_correction_callback;
		std::function<void(double)> post_grid_to_particle_transfer_callback;

		pcg32 random;
		std::vector<std::unique_ptr<source>> sources;
		vec3d grid_offset, gravity;
		double
			cfl_number = 3.0,
			blending_factor = 1.0,
			cell_size = std::numeric_limits<double>::quiet_NaN(),
			density = 1.0,
			boundary_skin_width = 0.1,
			correction_stiffness = 5.0;
		std::size_t velocity_extrapolation_iterations = 1;
		method simulation_method = method::apic;

		// ---- additions of the B200 build (not in the reference) ----
		double pressure_tolerance = 1e-6;          ///< pressure_solver::tolerance used inside time_step
		std::size_t pressure_max_iterations = 200; ///< pressure_solver::max_iterations used inside time_step
		preconditioner pressure_preconditioner = preconditioner::multigrid;
		int device = 0;                            ///< CUDA ordinal, read when the device context is created
		/// Sources (seed_cell, velocity coercion) run on the device inside the fused step.  Counts, cells and velocities
		/// of the spawned particles are the reference's; their positions come from a counter-based generator instead of
		/// \ref random.  Set to \p false to spawn on the host from \ref random exactly like the reference (staged step).
		bool device_sources = true;
		/// Residual / iteration count of the last pressure solve (what post_pressure_solve_callback receives).
		double last_residual = 0.0;
		std::size_t last_iterations = 0;
		/// The device context (created on first use); exposed for interop with other lfk consumers.
		lfk_ctx *device_context();
		/// Makes the host mirrors current without marking the device copies stale.
		void sync_to_host() const;
	private:
		struct _cell_particles {
			std::size_t begin = 0, count = 0;
		};
		mutable std::vector<particle> _particles;
		mutable mac_grid _grid, _old_grid;
		mutable grid3<_cell_particles> _space_hash; // host copy of the sorted-cell table (valid when _hash_host)
		mutable std::vector<std::size_t> _fluid_cells;

		mutable lfk_ctx *_ctx = nullptr;
		// coherence: which side holds the current particles / cells
		mutable bool _p_host = true, _p_dev = false, _g_host = true, _g_dev = false;
		mutable bool _hash_host = false, _hash_dev = false;
		bool _sources_dev = false; // the device holds a non-empty source list
		vec3s _size;

		void _ensure_ctx() const;
		void _push_params() const;
		void _particles_to_device() const;
		void _particles_to_host() const;
		void _grid_to_device() const;
		void _grid_to_host() const;
		void _table_to_host() const;
		void _check(int rc) const;
		bool _needs_staged_step() const;
		void _staged_time_step(double dt);
		void _update_sources();
		void _push_sources();
		void _coerce_source_velocities();
	};
}
