// Minimal stand-in so the reference's utils.cpp compiles without Boost (absent in this image).
// Only generateRandomHash() (sources/utils/utils.cpp:24-29) uses it; that function is off the hot path.
#pragma once
#include <array>
#include <cstdint>
#include <random>
#include <string>
namespace boost { namespace uuids {
struct uuid { std::array<uint8_t, 16> data{}; };
struct random_generator {
  uuid operator()() { std::random_device rd; uuid u; for (auto& b : u.data) b = static_cast<uint8_t>(rd()); return u; }
};
}}
