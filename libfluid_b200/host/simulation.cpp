// Host side of fluid::simulation / fluid::mac_grid / fluid::pressure_solver above the lfk C ABI.
// Orchestration follows the reference's time_step (src/simulation.cpp:43-125); every numerical stage is a device
// call.  There is no CPU fallback: a failing lfk call throws std::runtime_error with the library's message.
#include "fluid/simulation.h"

#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <string>

#include "fluid/pressure_solver.h"
#include "lfk.h"

namespace fluid {
	// ------------------------------------------------------------------------------------------ mac_grid
	std::pair<mac_grid::face_samples, vec3d> mac_grid::get_face_samples(vec3s gi, vec3d t) const {
		// host restatement of the sampling rule (reference src/mac_grid.cpp:40-112): cells gi-1..gi+1 per axis with
		// indices clamped into the grid; a component is read as 0 where its own axis was clamped or sits on the last
		// layer; the lower sample per axis is chosen by the half of the cell the point lies in
		const vec3s n = _grid.get_size();
		std::size_t ci[3][3];
		bool cl[3][3];
		for (int a = 0; a < 3; ++a) {
			for (std::size_t d = 0; d < 3; ++d) {
				std::size_t v = gi[a] + d;
				cl[a][d] = v < 1 || v >= n[a];
				ci[a][d] = (v < 1 ? 1 : (v >= n[a] ? n[a] : v)) - 1;
			}
		}
		vec3d tmid = t - vec3d(0.5, 0.5, 0.5);
		std::size_t sel[3] = { 1, 1, 1 };
		for (int a = 0; a < 3; ++a) {
			if (tmid[a] < 0.0) {
				sel[a] = 0;
				tmid[a] += 1.0;
			}
		}
		auto comp = [&](int k, std::size_t dx, std::size_t dy, std::size_t dz) {
			bool clamped = k == 0 ? cl[0][dx] : (k == 1 ? cl[1][dy] : cl[2][dz]);
			return clamped ? 0.0 : _grid(ci[0][dx], ci[1][dy], ci[2][dz]).velocities_posface[k];
		};
		face_samples r;
		vec3d *out[8] = { &r.v000, &r.v001, &r.v010, &r.v011, &r.v100, &r.v101, &r.v110, &r.v111 };
		for (std::size_t k = 0; k < 8; ++k) {
			std::size_t bx = k & 1, by = (k >> 1) & 1, bz = (k >> 2) & 1;
			*out[k] = vec3d(comp(0, bx, sel[1] + by, sel[2] + bz), comp(1, sel[0] + bx, by, sel[2] + bz),
				comp(2, sel[0] + bx, sel[1] + by, bz));
		}
		return { r, tmid };
	}

	// ---------------------------------------------------------------------------------------- simulation
	simulation::simulation() = default;
	simulation::~simulation() {
		if (_ctx) {
			lfk_destroy(_ctx);
		}
	}

	void simulation::_check(int rc) const {
		if (rc != 0) {
			throw std::runtime_error(std::string("libfluid_b200: ") + lfk_last_error(_ctx));
		}
	}

	void simulation::_ensure_ctx() const {
		if (_ctx) {
			return;
		}
		if (_size.x == 0 || _size.y == 0 || _size.z == 0) {
			throw std::runtime_error("libfluid_b200: simulation::resize() has not been called");
		}
		lfk_ctx *c = nullptr;
		int rc = lfk_create(&c, _size.x, _size.y, _size.z, device, nullptr, 1, 0, nullptr);
		if (rc != 0) {
			throw std::runtime_error(std::string("libfluid_b200: ") + lfk_last_error(nullptr));
		}
		_ctx = c;
	}

	void simulation::_push_params() const {
		_ensure_ctx();
		lfk_params p{};
		for (int d = 0; d < 3; ++d) {
			p.grid_offset[d] = grid_offset[d];
			p.gravity[d] = gravity[d];
		}
		p.cell_size = cell_size;
		p.density = density;
		p.boundary_skin_width = boundary_skin_width;
		p.correction_stiffness = correction_stiffness;
		p.blending_factor = blending_factor;
		p.cfl_number = cfl_number;
		p.tolerance = pressure_tolerance;
		p.method = static_cast<int>(simulation_method);
		p.extrapolation_iterations = static_cast<int>(velocity_extrapolation_iterations);
		p.max_iterations = static_cast<int>(pressure_max_iterations);
		p.preconditioner = pressure_preconditioner == preconditioner::multigrid ? LFK_PRECOND_MULTIGRID : LFK_PRECOND_JACOBI;
		_check(lfk_set_params(_ctx, &p));
	}

	lfk_ctx *simulation::device_context() {
		_push_params();
		return _ctx;
	}

	// ---- coherence of the host mirrors ----
	void simulation::_particles_to_device() const {
		_push_params();
		if (!_p_dev) {
			_check(lfk_upload_particles(_ctx, _particles.data(), _particles.size()));
			_p_dev = true;
			_hash_dev = false;
		}
	}
	void simulation::_particles_to_host() const {
		if (!_p_host) {
			std::uint64_t n = 0;
			_check(lfk_num_particles(_ctx, &n));
			_particles.resize(n);
			_check(lfk_download_particles(_ctx, _particles.data(), n, &n));
			_p_host = true;
		}
	}
	void simulation::_grid_to_device() const {
		_push_params();
		if (!_g_dev) {
			_check(lfk_upload_cells(_ctx, _grid.grid().data()));
			if (simulation_method == method::flip_blend && _old_grid.grid().get_size() == _size) {
				_check(lfk_upload_old_cells(_ctx, _old_grid.grid().data()));
			}
			_g_dev = true;
		}
	}
	void simulation::_grid_to_host() const {
		if (!_g_host) {
			_check(lfk_download_cells(_ctx, _grid.grid().data()));
			_g_host = true;
		}
	}
	void simulation::_table_to_host() const {
		if (_hash_host) {
			return;
		}
		std::size_t nc = _size.x * _size.y * _size.z;
		std::vector<std::uint64_t> b(nc), c(nc);
		_check(lfk_download_table(_ctx, b.data(), c.data()));
		for (std::size_t i = 0; i < nc; ++i) {
			_space_hash[i].begin = b[i];
			_space_hash[i].count = c[i];
		}
		std::uint64_t nf = 0;
		_check(lfk_num_fluid_cells(_ctx, &nf));
		std::vector<std::uint64_t> fc(nf);
		_check(lfk_download_fluid_cells(_ctx, fc.data(), nf));
		_fluid_cells.assign(fc.begin(), fc.end());
		_hash_host = true;
	}
	void simulation::sync_to_host() const {
		_particles_to_host();
		_grid_to_host();
	}

	mac_grid &simulation::grid() {
		_grid_to_host();
		_g_dev = false; // the caller may write cell types / velocities
		return _grid;
	}
	const mac_grid &simulation::grid() const {
		_grid_to_host();
		return _grid;
	}
	std::vector<simulation::particle> &simulation::particles() {
		_particles_to_host();
		_p_dev = false;
		_hash_dev = false;
		_hash_host = false;
		return _particles;
	}
	const std::vector<simulation::particle> &simulation::particles() const {
		_particles_to_host();
		return _particles;
	}

	void simulation::resize(vec3s sz) {
		if (_ctx) {
			lfk_destroy(_ctx);
			_ctx = nullptr;
		}
		_size = sz;
		_grid = mac_grid(sz);
		_old_grid = mac_grid();
		_space_hash = grid3<_cell_particles>(sz);
		_fluid_cells.clear();
		_p_host = _g_host = true;
		_p_dev = _g_dev = _hash_dev = _hash_host = false;
	}

	// ---- stepping ----
	void simulation::update(double dt) { // reference src/simulation.cpp:31-41
		while (true) {
			double ts = cfl_number * cfl();
			if (ts > dt) {
				time_step(dt);
				break;
			}
			time_step(ts);
			dt -= ts;
		}
	}
	void simulation::time_step() { // :127-129
		time_step(std::min(cfl_number * cfl(), 0.033));
	}

	bool simulation::_needs_staged_step() const {
		if (pre_time_step_callback || post_advection_callback || post_particle_to_grid_transfer_callback ||
			post_gravity_callback || post_pressure_solve_callback || post_apply_pressure_callback ||
			post_correction_callback) {
			return true; // post_grid_to_particle_transfer_callback fires after the step: the fused path can serve it
		}
		if (!device_sources) {
			for (const auto &s : sources) {
				if (s && s->active) {
					return true;
				}
			}
		}
		return false;
	}

	// simulation::sources -> lfk_set_sources (the list is small; it is pushed before every fused step because callers
	// edit `sources`, `active` and `velocity` freely between steps, testbed/main.cpp:156-165)
	void simulation::_push_sources() {
		std::vector<lfk_source> list;
		std::vector<std::vector<std::uint64_t>> cells;
		bool any = false;
		for (const auto &s : sources) {
			lfk_source d{};
			if (s) {
				cells.emplace_back();
				for (vec3s c : s->cells) {
					cells.back().push_back(c.x);
					cells.back().push_back(c.y);
					cells.back().push_back(c.z);
				}
				d.cells = cells.back().data();
				d.num_cells = s->cells.size();
				for (int k = 0; k < 3; ++k) {
					d.velocity[k] = s->velocity[k];
				}
				d.target_density_cubic_root = static_cast<std::uint32_t>(s->target_density_cubic_root);
				d.active = s->active ? 1 : 0;
				d.coerce_velocity = s->coerce_velocity ? 1 : 0;
				any |= s->active;
			}
			list.push_back(d);
		}
		if (!any && !_sources_dev) {
			return;
		}
		_check(lfk_set_sources(_ctx, list.data(), any ? list.size() : 0));
		_sources_dev = any;
	}

	void simulation::time_step(double dt) {
		if (_needs_staged_step()) {
			_staged_time_step(dt);
		} else {
			_particles_to_device();
			_grid_to_device();
			_push_sources();
			_check(lfk_time_step(_ctx, dt));
			_p_host = _g_host = false;
			_hash_host = false;
			_hash_dev = true;
			lfk_stats st;
			_check(lfk_get_stats(_ctx, &st));
			last_residual = st.pcg_residual;
			last_iterations = st.pcg_iterations;
		}
		if (post_grid_to_particle_transfer_callback) {
			post_grid_to_particle_transfer_callback(dt);
		}
	}

	// the reference's sequence with every callback point honoured (src/simulation.cpp:43-125)
	void simulation::_staged_time_step(double dt) {
		auto dev = [this]() {
			_particles_to_device();
			_grid_to_device();
		};
		auto dirty = [this]() {
			_p_host = _g_host = false;
		};
		auto table = [this]() { // a callback that touched particles() invalidates the sorted-cell table
			if (!_hash_dev) {
				_check(lfk_hash(_ctx));
				_hash_dev = true;
				_hash_host = false;
				_p_host = false;
			}
		};
		if (pre_time_step_callback) {
			pre_time_step_callback(dt);
		}
		bool have_sources = false;
		for (const auto &s : sources) {
			have_sources |= s && s->active;
		}
		if (have_sources) { // :49 + :227-238: the first hash only feeds source velocity coercion
			dev();
			_check(lfk_hash(_ctx));
			dirty();
			_hash_dev = true;
			_hash_host = false;
			_coerce_source_velocities();
		}
		dev();
		_check(lfk_advect(_ctx, dt));
		dirty();
		if (post_advection_callback) {
			post_advection_callback(dt);
		}
		dev();
		_check(lfk_collide(_ctx));
		_check(lfk_hash(_ctx)); // :62
		dirty();
		_hash_dev = true;
		_hash_host = false;
		if (have_sources) { // :63-64
			_update_sources();
			dev();
			_check(lfk_hash(_ctx));
			dirty();
			_hash_dev = true;
			_hash_host = false;
		}
		table();
		_check(lfk_p2g(_ctx));
		dirty();
		if (post_particle_to_grid_transfer_callback) {
			post_particle_to_grid_transfer_callback(dt);
		}
		dev();
		_check(lfk_gravity(_ctx, dt));
		dirty();
		if (post_gravity_callback) {
			post_gravity_callback(dt);
		}
		dev();
		table();
		double res = 0.0;
		std::uint64_t iters = 0;
		_check(lfk_pressure_solve(_ctx, dt, &res, &iters));
		last_residual = res;
		last_iterations = iters;
		if (post_pressure_solve_callback) {
			std::uint64_t nf = 0;
			_check(lfk_num_fluid_cells(_ctx, &nf));
			std::vector<double> pressure(nf);
			_check(lfk_download_pressure(_ctx, pressure.data(), nf));
			std::vector<double> before = pressure;
			post_pressure_solve_callback(dt, pressure, res, iters);
			if (pressure != before) { // the callback receives a mutable reference in the reference API
				_check(lfk_upload_pressure(_ctx, pressure.data(), nf));
			}
		}
		_check(lfk_apply_pressure(_ctx, dt));
		dirty();
		if (post_apply_pressure_callback) {
			post_apply_pressure_callback(dt);
		}
		dev();
		table();
		_check(lfk_correct(_ctx, dt));
		dirty();
		if (post_correction_callback) {
			post_correction_callback(dt);
		}
		dev();
		_check(lfk_collide(_ctx));
		// extrapolation reads the particle counts of the step's table (positions moved since, counts did not)
		_check(lfk_extrapolate(_ctx));
		_check(lfk_g2p(_ctx));
		dirty();
	}

	void simulation::_coerce_source_velocities() { // reference src/simulation.cpp:227-238
		bool any = false;
		for (const auto &s : sources) {
			any |= s && s->active && s->coerce_velocity;
		}
		if (!any) {
			return;
		}
		_table_to_host();
		std::vector<particle> &ps = particles();
		_hash_host = true; // particles() invalidated it, but nothing has been modified yet
		for (const auto &s : sources) {
			if (!s || !s->active || !s->coerce_velocity) {
				continue;
			}
			for (vec3s c : s->cells) {
				_cell_particles cp = _space_hash(c);
				for (std::size_t k = 0; k < cp.count; ++k) {
					particle &p = ps[cp.begin + k];
					p.velocity = s->velocity;
					p.cx = p.cy = p.cz = vec3d();
				}
			}
		}
	}

	void simulation::_update_sources() { // reference src/simulation.cpp:756-765
		_table_to_host();
		for (const auto &s : sources) {
			if (!s || !s->active) {
				continue;
			}
			for (vec3s c : s->cells) {
				seed_cell(c, s->velocity, s->target_density_cubic_root);
			}
		}
	}

	void simulation::reset_space_hash() {
		_space_hash.fill(_cell_particles());
		_fluid_cells.clear();
		_hash_host = true;
		_hash_dev = false;
	}
	void simulation::update_and_hash_particles() {
		_particles_to_device();
		_grid_to_device();
		_check(lfk_hash(_ctx));
		_p_host = false;
		_hash_dev = true;
		_hash_host = false;
	}
	void simulation::hash_particles() {
		// the reference sorts by the stored raw_cell_index; on the device the keys are recomputed from the positions,
		// which is what every caller (time_step, the Maya node after re-injecting particles) relies on anyway
		update_and_hash_particles();
	}

	void simulation::seed_cell(vec3s cell, vec3d velocity, std::size_t dens) { // reference src/simulation.cpp:136-151
		if (!_hash_host) {
			_table_to_host();
		}
		std::vector<particle> &ps = particles(); // invalidates the table flags; the host copy stays usable here
		_hash_host = true;
		std::size_t index = _grid.grid().index_to_raw(cell), num = _space_hash(cell).count, target = dens * dens * dens;
		std::uniform_real_distribution<double> dist(0.0, cell_size);
		vec3d offset = grid_offset + vec3d(cell) * cell_size;
		for (; num < target; ++num) {
			double jz = dist(random), jy = dist(random), jx = dist(random); // g++ argument order, see seed_func
			particle p;
			p.old_position = p.position = offset + vec3d(jx, jy, jz);
			p.velocity = velocity;
			p.raw_cell_index = index;
			ps.emplace_back(p);
		}
		_space_hash(cell).count = target;
	}

	void simulation::seed_box(vec3d start, vec3d size, vec3d vel, std::size_t dens) { // :153-167
		vec3d end = start + size;
		vec3s sc = world_position_to_cell_index_unclamped(start), ec = world_position_to_cell_index_unclamped(end);
		seed_func(
			sc, ec - sc + vec3s(1, 1, 1),
			[&](vec3d p) {
				return p.x > start.x && p.y > start.y && p.z > start.z && p.x < end.x && p.y < end.y && p.z < end.z;
			},
			vel, dens
		);
	}
	void simulation::seed_sphere(vec3d center, double radius, vec3d vel, std::size_t dens) { // :169-181
		vec3d r(radius, radius, radius);
		vec3s sc = world_position_to_cell_index_unclamped(center - r), ec = world_position_to_cell_index_unclamped(center + r);
		double r2 = radius * radius;
		seed_func(
			sc, ec - sc + vec3s(1, 1, 1),
			[&](vec3d p) {
				return (p - center).squared_length() < r2;
			},
			vel, dens
		);
	}

	vec3s simulation::world_position_to_cell_index_unclamped(vec3d pos) const { // :190-197
		vec3d g = (pos - grid_offset) / cell_size;
		vec3s r;
		for (std::size_t d = 0; d < 3; ++d) {
			r[d] = static_cast<std::size_t>(std::max(g[d], 0.0));
		}
		return r;
	}
	vec3s simulation::world_position_to_cell_index(vec3d pos) const { // :183-188
		vec3s r = world_position_to_cell_index_unclamped(pos);
		for (std::size_t d = 0; d < 3; ++d) {
			r[d] = std::min(r[d], _size[d]);
		}
		return r;
	}

	double simulation::cfl() const { // :199-205
		if (_p_dev && _ctx) {
			double v = 0.0;
			_check(lfk_cfl(_ctx, &v));
			return v;
		}
		double maxlen = 0.0;
		for (const particle &p : _particles) {
			maxlen = std::max(maxlen, p.velocity.squared_length());
		}
		return cell_size / std::sqrt(maxlen);
	}

	// ------------------------------------------------------------------------------------- pressure_solver
	pressure_solver::pressure_solver(simulation &sim, const std::vector<vec3s> &fluid_cells) :
		_fluid_cells(fluid_cells), _sim(sim) {
	}

	std::tuple<std::vector<double>, double, std::size_t> pressure_solver::solve(double dt) {
		_sim.pressure_tolerance = tolerance;
		_sim.pressure_max_iterations = max_iterations;
		_sim._particles_to_device();
		_sim._grid_to_device();
		if (!_sim._hash_dev) {
			_sim._check(lfk_hash(_sim._ctx));
			_sim._hash_dev = true;
			_sim._p_host = false;
		}
		double res = 0.0;
		std::uint64_t iters = 0, nf = 0;
		_sim._check(lfk_pressure_solve(_sim._ctx, dt, &res, &iters));
		_sim._check(lfk_num_fluid_cells(_sim._ctx, &nf));
		if (nf != _fluid_cells.size()) {
			throw std::runtime_error("libfluid_b200: fluid_cells does not match the cells that hold particles");
		}
		std::vector<double> p(nf);
		_sim._check(lfk_download_pressure(_sim._ctx, p.data(), nf));
		_sim.last_residual = res;
		_sim.last_iterations = iters;
		return { std::move(p), res, static_cast<std::size_t>(iters) };
	}

	void pressure_solver::apply_pressure(double dt, const std::vector<double> &pressure) const {
		_sim._particles_to_device();
		_sim._grid_to_device();
		if (!_sim._hash_dev) {
			_sim._check(lfk_hash(_sim._ctx));
			_sim._hash_dev = true;
			_sim._p_host = false;
		}
		_sim._check(lfk_download_rhs(_sim._ctx, dt, nullptr, nullptr, pressure.size())); // (re)builds the matrix flags
		_sim._check(lfk_upload_pressure(_sim._ctx, pressure.data(), pressure.size()));
		_sim._check(lfk_apply_pressure(_sim._ctx, dt));
		_sim._g_host = false;
	}
}
