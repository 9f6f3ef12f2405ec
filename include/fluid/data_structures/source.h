#pragma once
// fluid::source -- a fluid source (reference include/fluid/data_structures/source.h:12-22).
#include <vector>

#include "../math/vec.h"

namespace fluid {
	class source {
	public:
		std::vector<vec3s> cells;                  ///< cells in which particles are spawned
		vec3d velocity;                            ///< velocity of spawned particles
		std::size_t target_density_cubic_root = 2; ///< cubic root of the seeding density
		bool active = true;                        ///< whether the source is active
		bool coerce_velocity = false;              ///< override the velocity of every particle inside `cells`
	};
}
