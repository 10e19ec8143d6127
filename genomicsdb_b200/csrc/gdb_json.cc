#include "gdb_json.h"

#include <cstdio>
#include <cstdlib>
#include <stdexcept>

namespace gdb {

namespace {
struct Parser {
  const std::string& s; size_t p = 0;
  explicit Parser(const std::string& t) : s(t) {}
  [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("JSON: ") + m + " at offset " + std::to_string(p)); }
  void ws() { while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\n' || s[p] == '\r')) ++p; }
  bool lit(const char* w) { size_t n = strlen_(w); if (s.compare(p, n, w) == 0) { p += n; return true; } return false; }
  static size_t strlen_(const char* w) { size_t n = 0; while (w[n]) ++n; return n; }
  std::string str() {
    if (s[p] != '"') fail("expected string");
    ++p; std::string o;
    while (p < s.size() && s[p] != '"') {
      if (s[p] == '\\') {
        ++p; if (p >= s.size()) fail("bad escape");
        char c = s[p++];
        switch (c) {
          case 'n': o += '\n'; break; case 't': o += '\t'; break; case 'r': o += '\r'; break;
          case 'b': o += '\b'; break; case 'f': o += '\f'; break;
          case 'u': { if (p + 4 > s.size()) fail("bad \\u"); unsigned v = (unsigned)strtoul(s.substr(p, 4).c_str(), nullptr, 16); p += 4; if (v < 0x80) o += (char)v; else { o += (char)(0xC0 | (v >> 6)); o += (char)(0x80 | (v & 0x3F)); } break; }
          default: o += c;
        }
      } else o += s[p++];
    }
    if (p >= s.size()) fail("unterminated string");
    ++p; return o;
  }
  JsonValue value() {
    ws(); if (p >= s.size()) fail("unexpected end");
    JsonValue v;
    char c = s[p];
    if (c == '{') {
      v.type = JsonValue::OBJ; ++p; ws();
      if (s[p] == '}') { ++p; return v; }
      for (;;) { ws(); std::string k = str(); ws(); if (s[p] != ':') fail("expected ':'"); ++p; v.obj.emplace_back(k, value()); ws(); if (s[p] == ',') { ++p; continue; } if (s[p] == '}') { ++p; break; } fail("expected ',' or '}'"); }
    } else if (c == '[') {
      v.type = JsonValue::ARR; ++p; ws();
      if (s[p] == ']') { ++p; return v; }
      for (;;) { v.arr.push_back(value()); ws(); if (s[p] == ',') { ++p; continue; } if (s[p] == ']') { ++p; break; } fail("expected ',' or ']'"); }
    } else if (c == '"') { v.type = JsonValue::STR; v.str = str(); }
    else if (lit("true")) { v.type = JsonValue::BOOL; v.b = true; }
    else if (lit("false")) { v.type = JsonValue::BOOL; v.b = false; }
    else if (lit("null")) { v.type = JsonValue::NUL; }
    else {
      size_t q = p; bool is_int = true;
      if (s[q] == '-' || s[q] == '+') ++q;
      while (q < s.size() && ((s[q] >= '0' && s[q] <= '9') || s[q] == '.' || s[q] == 'e' || s[q] == 'E' || s[q] == '-' || s[q] == '+')) { if (s[q] == '.' || s[q] == 'e' || s[q] == 'E') is_int = false; ++q; }
      if (q == p) fail("unexpected character");
      std::string t = s.substr(p, q - p);
      v.type = JsonValue::NUM; v.is_int = is_int; v.num = strtod(t.c_str(), nullptr); if (is_int) v.inum = strtoll(t.c_str(), nullptr, 10);
      p = q;
    }
    return v;
  }
};
}  // namespace

const JsonValue& JsonValue::at(const std::string& k) const {
  const JsonValue* v = find(k);
  if (!v) throw std::runtime_error("JSON: missing key " + k);
  return *v;
}

JsonValue json_parse(const std::string& text) { Parser ps(text); JsonValue v = ps.value(); ps.ws(); return v; }

JsonValue json_parse_file(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open " + path);
  std::string t; char buf[1 << 16]; size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) t.append(buf, n);
  fclose(f);
  return json_parse(t);
}

}  // namespace gdb
