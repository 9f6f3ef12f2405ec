// extern "C" hooks used by tests/ and the examples to drive the C++ API mirror (include/fluid/*.h) from Python.
#include <cstring>
#include <memory>

#include "fluid/pressure_solver.h"
#include "fluid/simulation.h"

using fluid::vec3d;
using fluid::vec3s;
using sim_t = fluid::simulation;

extern "C" {
	void *hapi_create(std::size_t nx, std::size_t ny, std::size_t nz, double h, const double *off, const double *g,
		int method, double blend) {
		auto *s = new sim_t();
		s->resize(vec3s(nx, ny, nz));
		s->cell_size = h;
		s->grid_offset = vec3d(off[0], off[1], off[2]);
		s->gravity = vec3d(g[0], g[1], g[2]);
		s->simulation_method = static_cast<sim_t::method>(method);
		s->blending_factor = blend;
		s->reset_space_hash();
		return s;
	}
	void hapi_destroy(void *p) {
		delete static_cast<sim_t*>(p);
	}
	void hapi_seed_box(void *p, const double *a, const double *b, std::size_t dens) {
		static_cast<sim_t*>(p)->seed_box(vec3d(a[0], a[1], a[2]), vec3d(b[0], b[1], b[2]), vec3d(), dens);
	}
	void hapi_seed_sphere(void *p, const double *c, double r, std::size_t dens) {
		static_cast<sim_t*>(p)->seed_sphere(vec3d(c[0], c[1], c[2]), r, vec3d(), dens);
	}
	void hapi_set_solid(void *p, const unsigned char *mask) {
		auto &g = static_cast<sim_t*>(p)->grid().grid();
		std::size_t n = g.get_array_size(g.get_size());
		for (std::size_t i = 0; i < n; ++i) {
			if (mask[i]) {
				g[i].cell_type = fluid::mac_grid::cell::type::solid;
			}
		}
	}
	void hapi_add_source(void *p, const std::size_t *cells, std::size_t n, const double *vel, int coerce) {
		auto src = std::make_unique<fluid::source>();
		for (std::size_t i = 0; i < n; ++i) {
			src->cells.emplace_back(cells[3 * i], cells[3 * i + 1], cells[3 * i + 2]);
		}
		src->velocity = vec3d(vel[0], vel[1], vel[2]);
		src->coerce_velocity = coerce != 0;
		static_cast<sim_t*>(p)->sources.emplace_back(std::move(src));
	}
	std::size_t hapi_num_particles(void *p) {
		return static_cast<const sim_t*>(p)->particles().size();
	}
	void hapi_get_particles(void *p, void *out) {
		const auto &v = static_cast<const sim_t*>(p)->particles();
		std::memcpy(out, v.data(), v.size() * sizeof(sim_t::particle));
	}
	void hapi_set_particles(void *p, const void *in, std::size_t n) {
		auto &v = static_cast<sim_t*>(p)->particles();
		v.resize(n);
		std::memcpy(static_cast<void*>(v.data()), in, n * sizeof(sim_t::particle));
	}
	void hapi_get_cells(void *p, void *out) {
		const auto &g = static_cast<const sim_t*>(p)->grid().grid();
		std::memcpy(out, g.data(), g.get_array_size(g.get_size()) * sizeof(fluid::mac_grid::cell));
	}
	void hapi_set_cells(void *p, const void *in) {
		auto &g = static_cast<sim_t*>(p)->grid().grid();
		std::memcpy(static_cast<void*>(g.data()), in, g.get_array_size(g.get_size()) * sizeof(fluid::mac_grid::cell));
	}
	// returns 0, or -1 with the message in err (no CPU fallback: without a GPU this reports the lfk error)
	int hapi_time_step(void *p, double dt, char *err, std::size_t errlen) {
		try {
			if (dt > 0.0) {
				static_cast<sim_t*>(p)->time_step(dt);
			} else {
				static_cast<sim_t*>(p)->time_step();
			}
			return 0;
		} catch (const std::exception &e) {
			std::strncpy(err, e.what(), errlen - 1);
			err[errlen - 1] = 0;
			return -1;
		}
	}
	int hapi_update(void *p, double dt, char *err, std::size_t errlen) {
		try {
			static_cast<sim_t*>(p)->update(dt);
			return 0;
		} catch (const std::exception &e) {
			std::strncpy(err, e.what(), errlen - 1);
			err[errlen - 1] = 0;
			return -1;
		}
	}
	// installs every callback (forcing the staged path) and counts invocations; the pressure callback records
	// (residual, iterations, max pressure) like the testbed's diagnostics (testbed/main.cpp:104-116)
	struct hapi_cb_log {
		int calls[8];
		double residual, max_pressure;
		std::size_t iterations, pressure_len;
	};
	void hapi_install_callbacks(void *p, hapi_cb_log *log) {
		sim_t *s = static_cast<sim_t*>(p);
		std::memset(log, 0, sizeof(*log));
		s->pre_time_step_callback = [log](double) { ++log->calls[0]; };
		s->post_advection_callback = [log](double) { ++log->calls[1]; };
		s->post_particle_to_grid_transfer_callback = [log](double) { ++log->calls[2]; };
		s->post_gravity_callback = [log](double) { ++log->calls[3]; };
		s->post_pressure_solve_callback = [log](double, std::vector<double> &pr, double res, std::size_t it) {
			++log->calls[4];
			log->residual = res;
			log->iterations = it;
			log->pressure_len = pr.size();
			log->max_pressure = 0.0;
			for (double v : pr) {
				log->max_pressure = v > log->max_pressure ? v : log->max_pressure;
			}
		};
		s->post_apply_pressure_callback = [log](double) { ++log->calls[5]; };
		s->post_correction_callback = [log](double) { ++log->calls[6]; };
		s->post_grid_to_particle_transfer_callback = [log](double) { ++log->calls[7]; };
	}
	double hapi_cfl(void *p) {
		return static_cast<const sim_t*>(p)->cfl();
	}
	void hapi_last_solve(void *p, double *res, std::size_t *it) {
		*res = static_cast<sim_t*>(p)->last_residual;
		*it = static_cast<sim_t*>(p)->last_iterations;
	}
}
