// ORACLE / TEST INFRASTRUCTURE ONLY.  Stand-in for <boost/lexical_cast.hpp> (Boost >= 1.54,
// unpinned: reference README.md:10) so the reference TUs compile unmodified into oracle/_ref/.
// Restates lexical_cast<T>(std::string) for the arithmetic targets the reference uses
// (float, uint32_t): the WHOLE string must be consumed, no leading/trailing whitespace is
// accepted, failure throws boost::bad_lexical_cast.  Decimal -> float is correctly rounded
// (strtof), as any conforming parser is on the short decimals the reference's inputs use.
#ifndef VXB200_ORACLE_BOOST_LEXICAL_CAST_SHIM
#define VXB200_ORACLE_BOOST_LEXICAL_CAST_SHIM

#include <cctype>
#include <cerrno>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <typeinfo>

namespace boost {

class bad_lexical_cast : public std::bad_cast {
 public:
  const char* what() const noexcept override {
    return "bad lexical cast: source type value could not be interpreted as target";
  }
};

namespace shim_detail {
template <class T>
struct caster;

template <>
struct caster<float> {
  static float run(const std::string& s) {
    if (s.empty() || std::isspace(static_cast<unsigned char>(s[0]))) throw bad_lexical_cast();
    char* end = nullptr;
    const float v = std::strtof(s.c_str(), &end);
    if (end != s.c_str() + s.size()) throw bad_lexical_cast();
    return v;
  }
};

template <>
struct caster<double> {
  static double run(const std::string& s) {
    if (s.empty() || std::isspace(static_cast<unsigned char>(s[0]))) throw bad_lexical_cast();
    char* end = nullptr;
    const double v = std::strtod(s.c_str(), &end);
    if (end != s.c_str() + s.size()) throw bad_lexical_cast();
    return v;
  }
};

template <>
struct caster<uint32_t> {
  static uint32_t run(const std::string& s) {
    if (s.empty()) throw bad_lexical_cast();
    size_t i = 0;
    bool neg = false;
    if (s[0] == '+' || s[0] == '-') {
      neg = (s[0] == '-');
      i = 1;
    }
    if (i >= s.size()) throw bad_lexical_cast();
    uint64_t acc = 0;
    for (; i < s.size(); ++i) {
      if (s[i] < '0' || s[i] > '9') throw bad_lexical_cast();
      acc = acc * 10 + uint64_t(s[i] - '0');
      if (acc > 0xFFFFFFFFull) throw bad_lexical_cast();
    }
    return neg ? uint32_t(0u - uint32_t(acc)) : uint32_t(acc);
  }
};

template <>
struct caster<int32_t> {
  static int32_t run(const std::string& s) {
    if (s.empty() || std::isspace(static_cast<unsigned char>(s[0]))) throw bad_lexical_cast();
    char* end = nullptr;
    errno = 0;
    const long v = std::strtol(s.c_str(), &end, 10);
    if (end != s.c_str() + s.size() || errno != 0 || v > 2147483647L || v < -2147483648L) throw bad_lexical_cast();
    return int32_t(v);
  }
};
}  // namespace shim_detail

template <class T>
T lexical_cast(const std::string& s) {
  return shim_detail::caster<T>::run(s);
}

}  // namespace boost

#endif
