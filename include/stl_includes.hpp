/** stl_includes.hpp -- the standard headers the reference's sources pull in through this file
 *  (src/stl_includes.hpp:13-32). */
#ifndef STL_INCLUDES
#define STL_INCLUDES
#include <algorithm>
#include <bitset>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <iterator>
#include <random>
#include <stdexcept>
#include <unordered_map>
#include <utility>
#include <vector>
#if __cplusplus >= 202002L
#include <ranges>
#endif
#endif
