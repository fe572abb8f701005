#pragma once
#include <boost/uuid/uuid_generators.hpp>
#include <cstdio>
namespace boost { namespace uuids {
inline std::string to_string(const uuid& u) {
  char buf[37]; int p = 0;
  for (int i = 0; i < 16; ++i) { p += std::snprintf(buf + p, sizeof(buf) - p, "%02x", u.data[i]); if (i == 3 || i == 5 || i == 7 || i == 9) buf[p++] = '-'; }
  buf[p] = 0; return buf;
}
}}
