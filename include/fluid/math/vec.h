#pragma once
// fluid::vec<N, T> -- the small fixed-size vector type of libfluid's public API (reference
// include/fluid/math/vec.h), rewritten as a plain aggregate-like template.  Only what the simulation-facing API
// (fluid::simulation, fluid::mac_grid, fluid::pressure_solver, callbacks, testbed, Maya node) touches is provided:
// x/y/z members, indexing, memberwise arithmetic, scalar multiply/divide, dot, squared_length/length, conversion
// between element types, vec_ops::apply / for_each helpers, axis<I>().
#include <cmath>
#include <cstddef>
#include <type_traits>
#include <utility>

namespace fluid {
	template <std::size_t N, typename T> struct vec;

	namespace _vec_details {
		template <typename T> struct storage2 { T x{}, y{}; };
		template <typename T> struct storage3 { T x{}, y{}, z{}; };
		template <typename T> struct storage4 { T x{}, y{}, z{}, w{}; };
		template <std::size_t N, typename T> struct storage_n { T v[N]{}; };
		template <std::size_t N, typename T> using storage = std::conditional_t<
			N == 2, storage2<T>, std::conditional_t<N == 3, storage3<T>, std::conditional_t<N == 4, storage4<T>, storage_n<N, T>>>
		>;
	}

	template <std::size_t N, typename T> struct vec : _vec_details::storage<N, T> {
		static_assert(N >= 2, "vec needs at least two components");
		using value_type = T;
		constexpr static std::size_t dimensionality = N;

		vec() = default;
		template <typename... Args, typename = std::enable_if_t<sizeof...(Args) == N && (N > 1)>> vec(Args &&...args) {
			T tmp[N] = { static_cast<T>(std::forward<Args>(args))... };
			for (std::size_t i = 0; i < N; ++i) {
				(*this)[i] = tmp[i];
			}
		}
		template <typename U, typename = std::enable_if_t<!std::is_same_v<U, T>>> explicit vec(const vec<N, U> &o) {
			for (std::size_t i = 0; i < N; ++i) {
				(*this)[i] = static_cast<T>(o[i]); // truncation for double -> size_t, like the reference
			}
		}

		constexpr static std::size_t size() {
			return N;
		}
		T &at(std::size_t i) {
			return reinterpret_cast<T*>(this)[i];
		}
		T at(std::size_t i) const {
			return reinterpret_cast<const T*>(this)[i];
		}
		T &operator[](std::size_t i) {
			return at(i);
		}
		T operator[](std::size_t i) const {
			return at(i);
		}
		template <std::size_t I> static vec axis() {
			static_assert(I < N, "invalid axis");
			vec r;
			r[I] = static_cast<T>(1);
			return r;
		}

		vec &operator+=(const vec &r) { for (std::size_t i = 0; i < N; ++i) { at(i) += r[i]; } return *this; }
		vec &operator-=(const vec &r) { for (std::size_t i = 0; i < N; ++i) { at(i) -= r[i]; } return *this; }
		template <typename U> vec &operator*=(const U &s) { for (std::size_t i = 0; i < N; ++i) { at(i) *= s; } return *this; }
		template <typename U> vec &operator/=(const U &s) { for (std::size_t i = 0; i < N; ++i) { at(i) /= s; } return *this; }
		friend vec operator+(vec l, const vec &r) { return l += r; }
		friend vec operator-(vec l, const vec &r) { return l -= r; }
		friend vec operator-(vec l) { for (std::size_t i = 0; i < N; ++i) { l[i] = -l[i]; } return l; }
		template <typename U, typename = std::enable_if_t<std::is_arithmetic_v<U>>> friend vec operator*(vec l, const U &s) { return l *= s; }
		template <typename U, typename = std::enable_if_t<std::is_arithmetic_v<U>>> friend vec operator*(const U &s, vec r) { return r *= s; }
		template <typename U, typename = std::enable_if_t<std::is_arithmetic_v<U>>> friend vec operator/(vec l, const U &s) { return l /= s; }
		friend bool operator==(const vec &l, const vec &r) {
			for (std::size_t i = 0; i < N; ++i) {
				if (!(l[i] == r[i])) {
					return false;
				}
			}
			return true;
		}
		friend bool operator!=(const vec &l, const vec &r) { return !(l == r); }

		T squared_length() const {
			T r{};
			for (std::size_t i = 0; i < N; ++i) {
				r += at(i) * at(i);
			}
			return r;
		}
		T length() const {
			return std::sqrt(squared_length());
		}
	};

	namespace vec_ops {
		template <typename V> typename V::value_type dot(const V &a, const V &b) {
			typename V::value_type r{};
			for (std::size_t i = 0; i < V::size(); ++i) {
				r += a[i] * b[i];
			}
			return r;
		}
		template <typename Res, typename F, typename... Vs> Res apply(const F &f, const Vs &...vs) {
			Res r;
			for (std::size_t i = 0; i < Res::size(); ++i) {
				r[i] = f(vs[i]...);
			}
			return r;
		}
		template <typename Res, typename F, typename... Vs> void apply_to(Res &out, const F &f, const Vs &...vs) {
			for (std::size_t i = 0; i < Res::size(); ++i) {
				out[i] = f(vs[i]...);
			}
		}
		namespace memberwise {
			template <typename V> V mul(const V &a, const V &b) {
				V r;
				for (std::size_t i = 0; i < V::size(); ++i) {
					r[i] = a[i] * b[i];
				}
				return r;
			}
			template <typename V> V div(const V &a, const V &b) {
				V r;
				for (std::size_t i = 0; i < V::size(); ++i) {
					r[i] = a[i] / b[i];
				}
				return r;
			}
		}
	}

	using vec2d = vec<2, double>;
	using vec2s = vec<2, std::size_t>;
	using vec3d = vec<3, double>;
	using vec3f = vec<3, float>;
	using vec3s = vec<3, std::size_t>;
	using vec3i = vec<3, int>;
	using vec4d = vec<4, double>;
	static_assert(sizeof(vec3d) == 24 && sizeof(vec3s) == 24, "vec3 must be three packed scalars");
}
