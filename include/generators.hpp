/**
 * generators.hpp -- drop-in for the reference's pair generators (src/generators.hpp:20-58).  They fix
 * the ORDER of the pair lists, hence the layout of every result vector:
 *   generate_pairwise_from_vector : ring, (v[i], v[(i+1) mod n])
 *   generate_all_pairs_from_vector: all n*n ordered pairs incl. (i,i), row-major (i outer, j inner)
 */
#ifndef SKS_GENERATORS_HPP
#define SKS_GENERATORS_HPP
#include "stl_includes.hpp"

template <typename T>
std::pair<std::vector<T>, std::vector<T>> generate_pairwise_from_vector(const std::vector<T> &v)
{
    const size_t n = v.size();
    std::pair<std::vector<T>, std::vector<T>> out;
    out.first.reserve(n);
    out.second.reserve(n);
    for (size_t i = 0; i < n; ++i)
    {
        out.first.push_back(v[i]);
        out.second.push_back(v[(i + 1) % n]);
    }
    return out;
}

template <typename T>
std::pair<std::vector<T>, std::vector<T>> generate_all_pairs_from_vector(const std::vector<T> &v)
{
    const size_t n = v.size();
    std::pair<std::vector<T>, std::vector<T>> out;
    out.first.reserve(n * n);
    out.second.reserve(n * n);
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 0; j < n; ++j)
        {
            out.first.push_back(v[i]);
            out.second.push_back(v[j]);
        }
    return out;
}
#endif
