// ORACLE / TEST INFRASTRUCTURE ONLY.  Stand-in for glow's <glow/glutil.h> (jbehley/glow at an
// unpinned HEAD, fetched at configure time by the reference: external/glow.cmake:3-6).  The hot
// path only needs the type name glow::vec3 to exist (src/data/common.h:16).
#ifndef VXB200_ORACLE_GLOW_SHIM
#define VXB200_ORACLE_GLOW_SHIM
namespace glow {
struct vec3 {
  float x, y, z;
};
}  // namespace glow
#endif
