// ORACLE / TEST INFRASTRUCTURE ONLY.  Stand-in for <boost/tokenizer.hpp>: only
// boost::char_separator<char> + boost::tokenizer<char_separator<char>> over a std::string,
// as used by the reference's rv::split (src/rv/string_utils.cpp:41-51).  Restated semantics
// of char_separator with no kept delimiters:
//   keep_empty_tokens: every dropped delimiter ends a token; "" -> [] (token_iterator::initialize
//                      is never valid on an empty range), "a:" -> ["a",""],
//                      ":a" -> ["","a"], "a::b" -> ["a","","b"].
//   drop_empty_tokens: maximal runs of non-delimiter characters.
#ifndef VXB200_ORACLE_BOOST_TOKENIZER_SHIM
#define VXB200_ORACLE_BOOST_TOKENIZER_SHIM

#include <string>
#include <vector>

namespace boost {

enum empty_token_policy { drop_empty_tokens, keep_empty_tokens };

template <class Char>
class char_separator {
 public:
  explicit char_separator(const Char* dropped, const Char* kept = "", empty_token_policy p = drop_empty_tokens)
      : dropped_(dropped), kept_(kept), policy_(p) {}
  std::basic_string<Char> dropped_, kept_;
  empty_token_policy policy_;
};

template <class Sep>
class tokenizer {
 public:
  typedef std::vector<std::string>::const_iterator iterator;
  tokenizer(const std::string& s, const Sep& sep) {
    std::string cur;
    if (s.empty()) return;
    if (sep.policy_ == keep_empty_tokens) {
      for (size_t i = 0; i < s.size(); ++i) {
        if (sep.dropped_.find(s[i]) != std::string::npos) {
          tokens_.push_back(cur);
          cur.clear();
        } else {
          cur += s[i];
        }
      }
      tokens_.push_back(cur);
    } else {
      for (size_t i = 0; i < s.size(); ++i) {
        if (sep.dropped_.find(s[i]) != std::string::npos) {
          if (!cur.empty()) tokens_.push_back(cur);
          cur.clear();
        } else {
          cur += s[i];
        }
      }
      if (!cur.empty()) tokens_.push_back(cur);
    }
  }
  iterator begin() const { return tokens_.begin(); }
  iterator end() const { return tokens_.end(); }

 private:
  std::vector<std::string> tokens_;
};

}  // namespace boost

#endif
