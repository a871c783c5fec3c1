// TEST INFRASTRUCTURE (oracle). Stand-in for stdtensor's internal shape header, included by the
// reference at src/std_cuda_tensor.hpp:8-13 (rank_t, basic_shape<r> with .size()).
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>

namespace ttl
{
namespace internal
{
using rank_t = std::uint8_t;

template <rank_t r> struct basic_shape {
    std::array<int, r> dims;

    basic_shape() : dims{} {}

    template <typename... D> explicit basic_shape(D... d) : dims{{(int)d...}}
    {
        static_assert(sizeof...(D) == r, "rank mismatch");
    }

    std::size_t size() const
    {
        std::size_t n = 1;
        for (rank_t i = 0; i < r; ++i) n *= (std::size_t)dims[i];
        return n;
    }

    template <typename... I> std::size_t offset(I... i) const
    {
        static_assert(sizeof...(I) == r, "rank mismatch");
        const int idx[r] = {(int)i...};
        std::size_t off = 0;
        for (rank_t k = 0; k < r; ++k) off = off * dims[k] + idx[k];
        return off;
    }

    basic_shape<(rank_t)(r - 1)> drop_front() const
    {
        basic_shape<(rank_t)(r - 1)> s;
        for (rank_t k = 1; k < r; ++k) s.dims[k - 1] = dims[k];
        return s;
    }
};

template <> struct basic_shape<0> {
    std::array<int, 1> dims;  // unused
    std::size_t size() const { return 1; }
};
}  // namespace internal
}  // namespace ttl
