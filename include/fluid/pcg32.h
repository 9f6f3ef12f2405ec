#pragma once
// Minimal PCG32 (XSH-RR 64/32, the `pcg32` typedef of pcg-cpp) as a UniformRandomBitGenerator.  The reference
// seeds particles from a default-constructed pcg32 (include/fluid/simulation.h:101,177); this engine produces the
// same stream (algorithm: M. E. O'Neill, "PCG: A Family of Simple Fast Space-Efficient Statistically Good
// Algorithms for Random Number Generation", 2014), so host-side seeding stays bit-compatible without vendoring
// pcg-cpp.  tests/test_host_api.py checks it against the reference's seeding.
#include <cstdint>
#include <limits>

namespace fluid {
	class pcg32_engine {
	public:
		using result_type = std::uint32_t;
		constexpr static std::uint64_t multiplier = 6364136223846793005ull;
		constexpr static std::uint64_t default_increment = 1442695040888963407ull;
		constexpr static std::uint64_t default_seed = 0xcafef00dd15ea5e5ull;

		explicit pcg32_engine(std::uint64_t seed = default_seed) : _inc(default_increment) {
			_state = (seed + _inc) * multiplier + _inc;
		}
		pcg32_engine(std::uint64_t seed, std::uint64_t stream) : _inc((stream << 1) | 1u) {
			_state = (seed + _inc) * multiplier + _inc;
		}
		constexpr static result_type min() { return 0; }
		constexpr static result_type max() { return std::numeric_limits<result_type>::max(); }
		result_type operator()() {
			std::uint64_t old = _state;
			_state = old * multiplier + _inc;
			std::uint32_t xorshifted = static_cast<std::uint32_t>(((old >> 18u) ^ old) >> 27u);
			std::uint32_t rot = static_cast<std::uint32_t>(old >> 59u);
			return (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
		}
	private:
		std::uint64_t _state, _inc;
	};
}
using pcg32 = fluid::pcg32_engine; // the name the reference API exposes (simulation::random)
