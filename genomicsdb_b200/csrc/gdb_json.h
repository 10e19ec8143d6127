// Minimal JSON reader for the loader / query / vid-mapping files GenomicsDB uses (the reference reads them with RapidJSON:
// src/main/cpp/src/config/json_config.cc:69-818).  Host-side configuration only; nothing here is on the hot path.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace gdb {

struct JsonValue {
  enum Type { NUL, BOOL, NUM, STR, ARR, OBJ } type = NUL;
  bool b = false;
  double num = 0;
  int64_t inum = 0;
  bool is_int = false;
  std::string str;
  std::vector<JsonValue> arr;
  std::vector<std::pair<std::string, JsonValue>> obj;   // insertion order preserved (vid "fields" order matters)

  bool has(const std::string& k) const { return find(k) != nullptr; }
  const JsonValue* find(const std::string& k) const {
    for (auto& kv : obj) if (kv.first == k) return &kv.second;
    return nullptr;
  }
  const JsonValue& at(const std::string& k) const;
  int64_t as_int() const { return is_int ? inum : (int64_t)num; }
  bool as_bool() const { return type == BOOL ? b : (type == NUM ? num != 0 : false); }
};

JsonValue json_parse(const std::string& text);        // throws std::runtime_error
JsonValue json_parse_file(const std::string& path);   // throws std::runtime_error

}  // namespace gdb
