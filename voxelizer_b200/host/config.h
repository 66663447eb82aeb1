// Config + parseConfiguration: same names, fields and file semantics as the reference
// (src/data/voxelize_utils.h:8-31, voxelize_utils.cpp:9-171), including its quirks:
// keys are compared untrimmed, values are trimmed of SPACES only, "past distance" overwrites
// pastScans, unknown keys are reported on stdout, an unreadable file yields the defaults.
#ifndef VOXELIZER_B200_HOST_CONFIG_H_
#define VOXELIZER_B200_HOST_CONFIG_H_

#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace vxb {

struct Vector4f {
  float v[4];
  Vector4f() : v{0, 0, 0, 0} {}
  Vector4f(float a, float b, float c, float d) : v{a, b, c, d} {}
  float& operator[](int i) { return v[i]; }
  const float& operator[](int i) const { return v[i]; }
  float x() const { return v[0]; }
  float y() const { return v[1]; }
  float z() const { return v[2]; }
};

struct Vector3f {
  float v[3];
  Vector3f() : v{0, 0, 0} {}
  Vector3f(float a, float b, float c) : v{a, b, c} {}
  float& operator[](int i) { return v[i]; }
  const float& operator[](int i) const { return v[i]; }
  static Vector3f Zero() { return Vector3f(); }
};

class Config {
 public:
  float voxelSize{0.5f};
  Vector4f minExtent{0, -20, -2, 1};
  Vector4f maxExtent{40, 20, 1, 1};
  uint32_t maxNumScans{100};
  uint32_t priorScans{0};
  uint32_t pastScans{10};
  float pastDistance{0.0f};
  float minRange{2.5f};
  float maxRange{25.0f};
  std::vector<uint32_t> filteredLabels;
  std::map<uint32_t, std::vector<uint32_t>> joinedLabels;
  uint32_t stride_num{0};
  float stride_distance{0.0f};
  bool hidecar{true};
};

// thrown where the reference lets boost::bad_lexical_cast escape
struct bad_lexical_cast : std::runtime_error {
  bad_lexical_cast() : std::runtime_error("bad lexical cast: source type value could not be interpreted as target") {}
};

Config parseConfiguration(const std::string& filename);

// string helpers with the reference's behaviour (rv/string_utils.h:23-28, string_utils.cpp:10-51)
std::string trim(const std::string& s);                                             // strips ' ' only
std::vector<std::string> split(const std::string& line, const std::string& delim);  // keeps empty tokens
float lexical_float(const std::string& s);
uint32_t lexical_u32(const std::string& s);

}  // namespace vxb

#endif
