#pragma once
#include <cstdint>
#include <vector>
namespace dq { namespace diffhost { struct Streams { std::vector<uint8_t> ctrl, diff, extra; int64_t visits = 0; }; } }
