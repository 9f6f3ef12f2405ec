#pragma once
// fluid::mac_grid -- host mirror of the MAC grid with the reference's AoS cell layout (include/fluid/mac_grid.h:12-73).
// The authoritative copy lives on the device as SoA face arrays; fluid::simulation keeps this mirror coherent
// lazily (see include/fluid/simulation.h).
#include <utility>

#include "data_structures/grid.h"

namespace fluid {
	class mac_grid {
	public:
		struct cell {
			enum class type : unsigned char {
				air = 0x1,
				fluid = 0x2,
				solid = 0x4,
			};
			vec3d velocities_posface;    ///< velocities on the +x, +y, +z faces
			type cell_type = type::air;
		};
		struct face_samples {
			vec3d v000, v001, v010, v011, v100, v101, v110, v111;
		};
		static_assert(sizeof(cell) == 32, "lfk_upload_cells expects the reference's 32-byte cell records");

		mac_grid() = default;
		explicit mac_grid(vec3s size) : _grid(size) {
		}

		/// the 2x2x2 staggered samples around a position (host version of reference src/mac_grid.cpp:51-112)
		std::pair<face_samples, vec3d> get_face_samples(vec3s grid_index, vec3d offset) const;

		cell *get_cell(vec3s i) {
			vec3s s = _grid.get_size();
			return (i.x >= s.x || i.y >= s.y || i.z >= s.z) ? nullptr : &_grid(i);
		}
		const cell *get_cell(vec3s i) const {
			vec3s s = _grid.get_size();
			return (i.x >= s.x || i.y >= s.y || i.z >= s.z) ? nullptr : &_grid(i);
		}
		/// out-of-grid cells read as solid
		std::pair<cell*, cell::type> get_cell_and_type(vec3s i) {
			cell *c = get_cell(i);
			return { c, c ? c->cell_type : cell::type::solid };
		}
		std::pair<const cell*, cell::type> get_cell_and_type(vec3s i) const {
			const cell *c = get_cell(i);
			return { c, c ? c->cell_type : cell::type::solid };
		}

		grid3<cell> &grid() { return _grid; }
		const grid3<cell> &grid() const { return _grid; }
	protected:
		grid3<cell> _grid;
	};
}
