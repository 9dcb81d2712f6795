// params.h — tokenizer and typed parameter lists for the .pbrt front end.
// Host-side mirror of reference src/core/parser/{parser,common}.rs and
// src/core/param_set/{param_set,texture_params}.rs for the subset of directives the
// PathIntegrator hot path needs (see DESIGN.md "front end").
#pragma once
#include <cstdint>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace pth {

struct ParseError : std::runtime_error {
  explicit ParseError(const std::string& m) : std::runtime_error(m) {}
};

enum TokKind { TK_EOF, TK_IDENT, TK_STRING, TK_NUMBER, TK_LBRACK, TK_RBRACK };
struct Token {
  TokKind kind;
  std::string text;
};

class Lexer {
 public:
  explicit Lexer(const std::string& s) : s_(s), i_(0) {}
  Token next() {
    skip_ws();
    if (i_ >= s_.size()) return Token{TK_EOF, ""};
    char c = s_[i_];
    if (c == '[') {
      i_++;
      return Token{TK_LBRACK, "["};
    }
    if (c == ']') {
      i_++;
      return Token{TK_RBRACK, "]"};
    }
    if (c == '"') {
      size_t j = i_ + 1;
      while (j < s_.size() && s_[j] != '"') j++;
      if (j >= s_.size()) throw ParseError("unterminated string literal");
      Token t{TK_STRING, s_.substr(i_ + 1, j - i_ - 1)};
      i_ = j + 1;
      return t;
    }
    size_t j = i_;
    while (j < s_.size() && !isspace((unsigned char)s_[j]) && s_[j] != '[' && s_[j] != ']' && s_[j] != '"' && s_[j] != '#') j++;
    Token t;
    t.text = s_.substr(i_, j - i_);
    i_ = j;
    char f = t.text[0];
    t.kind = (isdigit((unsigned char)f) || f == '-' || f == '+' || f == '.') ? TK_NUMBER : TK_IDENT;
    return t;
  }
  Token peek() {
    size_t save = i_;
    Token t = next();
    i_ = save;
    return t;
  }

 private:
  void skip_ws() {
    for (;;) {
      while (i_ < s_.size() && isspace((unsigned char)s_[i_])) i_++;
      if (i_ < s_.size() && s_[i_] == '#') {
        while (i_ < s_.size() && s_[i_] != '\n') i_++;
        continue;
      }
      break;
    }
  }
  const std::string& s_;
  size_t i_;
};

// One parameter: `"type name" values`.  cls groups the reference's typed maps.
enum ParamClass { PC_FLOAT, PC_INT, PC_BOOL, PC_STRING, PC_TEXTURE, PC_RGB, PC_XYZ, PC_BLACKBODY, PC_SPECTRUM, PC_POINT };
struct Param {
  ParamClass cls;
  std::vector<float> f;        // float / rgb / xyz / blackbody / point
  std::vector<int32_t> i;      // int
  std::vector<uint8_t> b;      // bool
  std::vector<std::string> s;  // string / texture / spectrum(file or pairs)
  size_t len() const {
    switch (cls) {
      case PC_INT: return i.size();
      case PC_BOOL: return b.size();
      case PC_STRING: case PC_TEXTURE: case PC_SPECTRUM: return s.size();
      default: return f.size();
    }
  }
};

struct ParamSet {
  std::map<std::string, Param> p;  // keyed by name (the reference keeps one map per type; a name is
                                   // looked up in the map of the requested type only, so key = cls:name)
  static std::string key(ParamClass c, const std::string& n) { return std::to_string((int)c) + ":" + n; }
  const Param* find(ParamClass c, const std::string& n) const {
    auto it = p.find(key(c, n));
    return it == p.end() ? nullptr : &it->second;
  }
  void set(ParamClass c, const std::string& n, const Param& v) { p[key(c, n)] = v; }
  float one_float(const std::string& n, float d) const {
    const Param* q = find(PC_FLOAT, n);
    return (q && q->f.size() == 1) ? q->f[0] : d;
  }
  int one_int(const std::string& n, int d) const {
    const Param* q = find(PC_INT, n);
    return (q && q->i.size() == 1) ? q->i[0] : d;
  }
  bool one_bool(const std::string& n, bool d) const {
    const Param* q = find(PC_BOOL, n);
    return (q && q->b.size() == 1) ? (q->b[0] != 0) : d;
  }
  std::string one_string(const std::string& n, const std::string& d) const {
    const Param* q = find(PC_STRING, n);
    return (q && q->s.size() == 1) ? q->s[0] : d;
  }
  bool one_point(const std::string& n, float out[3]) const {
    const Param* q = find(PC_POINT, n);
    if (q && q->f.size() == 3) {
      out[0] = q->f[0]; out[1] = q->f[1]; out[2] = q->f[2];
      return true;
    }
    return false;
  }
};

// wellknown_params.rs: untyped keys get a default type
inline const char* wellknown_type(const std::string& k) {
  static const std::map<std::string, const char*> m = {
      {"xresolution", "integer"}, {"yresolution", "integer"}, {"pixelsamples", "integer"}, {"maxdepth", "integer"},
      {"indices", "integer"},     {"P", "point"},             {"N", "normal"},             {"S", "vector"},
      {"uv", "float"},            {"st", "float"},            {"fov", "float"},            {"filename", "string"},
      {"Kd", "rgb"},              {"Ks", "rgb"},              {"Kr", "rgb"},               {"Kt", "rgb"},
      {"L", "rgb"},               {"I", "rgb"},               {"roughness", "float"},      {"eta", "float"},
      {"from", "point"},          {"to", "point"},            {"scale", "rgb"},
  };
  auto it = m.find(k);
  return it == m.end() ? "" : it->second;
}

inline float parse_f32(const std::string& s) {  // Rust str::parse::<f32> is correctly rounded -> strtof
  char* end = nullptr;
  float v = strtof(s.c_str(), &end);
  if (end == s.c_str()) throw ParseError("bad number '" + s + "'");
  return v;
}

}  // namespace pth
