#pragma once
// fluid::grid<Dim, Cell> -- dense x-fastest cell array with the indexing convention of the reference
// (include/fluid/data_structures/grid.h:13-288): raw = x + nx * (y + ny * z).
#include <cassert>
#include <vector>

#include "../math/vec.h"

namespace fluid {
	template <std::size_t Dim, typename Cell> class grid {
	public:
		using size_type = vec<Dim, std::size_t>;

		grid() = default;
		explicit grid(size_type size) : grid(size, Cell{}) {
		}
		grid(size_type size, const Cell &c) : _cells(get_array_size(size), c), _size(size) {
			std::size_t stride = 1;
			for (std::size_t d = 0; d < Dim; ++d) {
				_stride[d] = stride;
				stride *= size[d];
			}
		}

		Cell &at(size_type i) { return _cells[index_to_raw(i)]; }
		const Cell &at(size_type i) const { return _cells[index_to_raw(i)]; }
		Cell &operator()(size_type i) { return at(i); }
		const Cell &operator()(size_type i) const { return at(i); }
		template <typename... A, typename = std::enable_if_t<sizeof...(A) == Dim && (Dim > 1)>> Cell &operator()(A... a) {
			return at(size_type(static_cast<std::size_t>(a)...));
		}
		template <typename... A, typename = std::enable_if_t<sizeof...(A) == Dim && (Dim > 1)>> const Cell &operator()(A... a) const {
			return at(size_type(static_cast<std::size_t>(a)...));
		}
		Cell &at_raw(std::size_t i) { return _cells[i]; }
		const Cell &at_raw(std::size_t i) const { return _cells[i]; }
		Cell &operator[](std::size_t i) { return _cells[i]; }
		const Cell &operator[](std::size_t i) const { return _cells[i]; }

		size_type get_size() const { return _size; }
		void fill(const Cell &value) {
			for (Cell &c : _cells) {
				c = value;
			}
		}
		bool is_border_cell(size_type i) const {
			for (std::size_t d = 0; d < Dim; ++d) {
				if (i[d] == 0 || i[d] + 1 == _size[d]) {
					return true;
				}
			}
			return false;
		}

		std::size_t index_to_raw(size_type i) const {
			std::size_t r = 0;
			for (std::size_t d = 0; d < Dim; ++d) {
				assert(i[d] < _size[d]);
				r += i[d] * _stride[d];
			}
			return r;
		}
		size_type index_from_raw(std::size_t raw) const {
			size_type r;
			for (std::size_t d = 0; d < Dim; ++d) {
				r[d] = raw % _size[d];
				raw /= _size[d];
			}
			return r;
		}
		static std::size_t get_array_size(size_type size) {
			std::size_t n = 1;
			for (std::size_t d = 0; d < Dim; ++d) {
				n *= size[d];
			}
			return n;
		}

		// callbacks receive (index, cell&), cells visited in memory order
		template <typename Cb> void for_each(Cb &&cb) {
			for_each_in_range_unchecked(std::forward<Cb>(cb), size_type(), _size);
		}
		template <typename Cb> void for_each_in_range_unchecked(Cb &&cb, size_type lo, size_type hi) {
			for (std::size_t d = 0; d < Dim; ++d) {
				if (lo[d] >= hi[d]) {
					return;
				}
			}
			size_type cur = lo;
			while (true) {
				cb(cur, _cells[index_to_raw(cur)]);
				std::size_t d = 0;
				for (; d < Dim; ++d) {
					if (++cur[d] < hi[d]) {
						break;
					}
					cur[d] = lo[d];
				}
				if (d == Dim) {
					return;
				}
			}
		}
		template <typename Cb> void for_each_in_range_checked(Cb &&cb, size_type lo, size_type hi) {
			for (std::size_t d = 0; d < Dim; ++d) {
				hi[d] = hi[d] < _size[d] ? hi[d] : _size[d];
			}
			for_each_in_range_unchecked(std::forward<Cb>(cb), lo, hi);
		}
		template <typename Cb> void for_each_in_range_checked(Cb &&cb, size_type center, size_type dmin, size_type dmax) {
			size_type lo, hi;
			for (std::size_t d = 0; d < Dim; ++d) {
				lo[d] = center[d] < dmin[d] ? 0 : center[d] - dmin[d];
				hi[d] = center[d] + dmax[d] + 1;
			}
			for_each_in_range_checked(std::forward<Cb>(cb), lo, hi);
		}

		Cell *data() { return _cells.data(); }
		const Cell *data() const { return _cells.data(); }
	private:
		std::vector<Cell> _cells;
		size_type _size, _stride;
	};
	template <typename Cell> using grid2 = grid<2, Cell>;
	template <typename Cell> using grid3 = grid<3, Cell>;
}
