#include "config.h"

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <iostream>

namespace vxb {

std::string trim(const std::string& s) {
  // The reference passes the literal " \0\t\n\r\x0B" as a std::string, which ends at the NUL: the
  // effective whitespace set is the single space (rv/string_utils.h:23-24).
  const size_t n = s.size();
  size_t first = 0;
  while (first < n && s[first] == ' ') ++first;
  if (first == n) return "";
  size_t last = n - 1;
  while (last > first && s[last] == ' ') --last;
  return s.substr(first, last - first + 1);
}

std::vector<std::string> split(const std::string& line, const std::string& delim) {
  // boost::tokenizer yields NO token for an empty range (token_iterator::initialize: valid_ = begin != end ? ... : false);
  // otherwise keep_empty_tokens closes a token at every delimiter ("a:" -> ["a", ""]).
  if (line.empty()) return {};
  std::vector<std::string> tokens(1);
  for (const char ch : line) {
    if (delim.find(ch) == std::string::npos) tokens.back().push_back(ch);
    else tokens.emplace_back();
  }
  return tokens;
}

float lexical_float(const std::string& s) {
  if (s.empty() || std::isspace(static_cast<unsigned char>(s.front()))) throw bad_lexical_cast();
  char* stop = nullptr;
  const float value = std::strtof(s.c_str(), &stop);
  if (stop != s.c_str() + s.size()) throw bad_lexical_cast();
  return value;
}

uint32_t lexical_u32(const std::string& s) {
  size_t pos = 0;
  bool negative = false;
  if (!s.empty() && (s[0] == '+' || s[0] == '-')) {
    negative = s[0] == '-';
    pos = 1;
  }
  if (pos >= s.size()) throw bad_lexical_cast();
  uint64_t value = 0;
  for (; pos < s.size(); ++pos) {
    const char ch = s[pos];
    if (ch < '0' || ch > '9') throw bad_lexical_cast();
    value = value * 10u + static_cast<uint64_t>(ch - '0');
    if (value > 0xFFFFFFFFull) throw bad_lexical_cast();
  }
  const uint32_t v32 = static_cast<uint32_t>(value);
  return negative ? static_cast<uint32_t>(0u - v32) : v32;
}

namespace {

bool bracketed(const std::string& s, char open, char close) { return !s.empty() && s.front() == open && s.back() == close; }

template <class T, class Conv>
std::vector<T> number_list(const std::string& text, Conv conv) {
  const std::string s = trim(text);
  std::vector<T> values;
  if (!bracketed(s, '[', ']')) {
    std::cerr << "Parser Error: " << s << " is not a valid list token!" << std::endl;
    return values;
  }
  for (const std::string& tok : split(s.substr(1, s.size() - 2), ",")) values.push_back(conv(trim(tok)));
  return values;
}

// "[{a: [b,c]}, {d: [e]}]" -> "{a: [b,c]}", " {d: [e]}" : pieces split at ',' are glued back
// together while a '[' is still open (voxelize_utils.cpp:43-72).
std::vector<std::string> item_list(const std::string& text) {
  const std::string s = trim(text);
  if (!bracketed(s, '[', ']')) std::cerr << "Parser Error: " << s << " is not a valid list token!" << std::endl;
  std::vector<std::string> items;
  if (s.size() < 2) return items;
  const std::vector<std::string> pieces = split(s.substr(1, s.size() - 2), ",");
  for (size_t i = 0; i < pieces.size(); ++i) {
    std::string item = pieces[i];
    const size_t open = item.find('[');
    if (open != std::string::npos && item.find(']', open) == std::string::npos) {
      for (;;) {
        if (i + 1 >= pieces.size()) throw std::out_of_range("unterminated list in configuration");
        const std::string& next = pieces[++i];
        item += "," + next;
        if (next.find(']') != std::string::npos) break;
      }
    }
    items.push_back(item);
  }
  return items;
}

std::vector<std::string> dictionary_entry(const std::string& text) {
  const std::string s = trim(text);
  if (!bracketed(s, '{', '}')) {
    std::cerr << "Parser Error: " << s << " is not a valid dictionary token!" << std::endl;
    return {};
  }
  return split(s.substr(1, s.size() - 2), ":");
}

Vector4f extent(const std::string& value) {
  const std::vector<float> c = number_list<float>(value, lexical_float);
  if (c.size() < 3) throw std::out_of_range("extent needs three coordinates");
  return Vector4f(c[0], c[1], c[2], 1.0f);
}

}  // namespace

Config parseConfiguration(const std::string& filename) {
  Config config;
  std::ifstream in(filename);
  if (!in.is_open()) return config;

  std::string line;
  in.peek();
  while (in.good() && !in.eof()) {
    std::getline(in, line);
    const std::string stripped = trim(line);
    if (!stripped.empty() && stripped[0] == '#') continue;

    const std::vector<std::string> parts = split(line, ":");
    if (parts.size() < 2) continue;
    const std::string& key = parts[0];
    std::string value = parts[1];
    for (size_t i = 2; i < parts.size(); ++i) value += ":" + parts[i];

    if (key == "max scans") config.maxNumScans = lexical_u32(trim(value));
    else if (key == "max range") config.maxRange = lexical_float(trim(value));
    else if (key == "voxel size") config.voxelSize = lexical_float(trim(value));
    else if (key == "min range") config.minRange = lexical_float(trim(value));
    else if (key == "prior scans") config.priorScans = lexical_u32(trim(value));
    else if (key == "past scans") config.pastScans = lexical_u32(trim(value));
    else if (key == "past distance") config.pastScans = static_cast<uint32_t>(lexical_float(trim(value)));  // as the reference
    else if (key == "stride num") config.stride_num = lexical_u32(trim(value));
    else if (key == "stride distance") config.stride_distance = lexical_float(trim(value));
    else if (key == "min extent") config.minExtent = extent(value);
    else if (key == "max extent") config.maxExtent = extent(value);
    else if (key == "ignore") config.filteredLabels = number_list<uint32_t>(value, lexical_u32);
    else if (key == "join") {
      for (const std::string& item : item_list(trim(value))) {
        const std::vector<std::string> kv = dictionary_entry(item);
        if (kv.size() < 2) throw std::out_of_range("join entry needs a key and a list");
        config.joinedLabels[lexical_u32(trim(kv[0]))] = number_list<uint32_t>(trim(kv[1]), lexical_u32);
      }
    } else {
      std::cout << "unknown parameter: " << key << std::endl;
    }
  }
  return config;
}

}  // namespace vxb
