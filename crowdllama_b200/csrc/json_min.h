// json_min.h — minimal JSON reader shared by the tokenizer.json loader (tokenizer.cpp) and the safetensors header
// reader (weights_io.cpp): objects, arrays, strings with escapes, numbers, true / false / null.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace cl {
namespace jsonmin {

// ================================================================================================
// minimal JSON reader (objects, arrays, strings with escapes, numbers, true / false / null)
// ================================================================================================

struct JVal {
  enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
  bool b = false;
  double n = 0;
  std::string s;
  std::vector<JVal> a;
  std::vector<std::pair<std::string, JVal>> o;
  const JVal* get(const char* key) const {
    if (type != Obj) return nullptr;
    for (auto& kv : o) if (kv.first == key) return &kv.second;
    return nullptr;
  }
  const std::string& str(const char* key, const std::string& dflt) const {
    const JVal* v = get(key);
    return v && v->type == Str ? v->s : dflt;
  }
  bool boolean(const char* key, bool dflt) const {
    const JVal* v = get(key);
    return v && v->type == Bool ? v->b : dflt;
  }
};

inline void append_utf8(std::string& out, uint32_t cp) {
  if (cp < 0x80) out.push_back((char)cp);
  else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
  else if (cp < 0x10000) { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
  else { out.push_back((char)(0xF0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 0x3F))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
}

struct JParser {
  const char* p; const char* end; std::string err;
  void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
  bool fail(const char* m) { if (err.empty()) err = m; return false; }
  bool hex4(uint32_t* v) {
    if (end - p < 4) return fail("truncated \\u escape");
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
      const char c = *p++;
      r = r * 16 + (c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : 99);
      if (r > 0xFFFFF) return fail("bad \\u escape");
    }
    *v = r;
    return true;
  }
  bool string(std::string* out) {
    if (p >= end || *p != '"') return fail("expected string");
    ++p;
    out->clear();
    while (p < end && *p != '"') {
      if (*p != '\\') { out->push_back(*p++); continue; }
      if (++p >= end) return fail("truncated escape");
      const char c = *p++;
      switch (c) {
        case '"': out->push_back('"'); break; case '\\': out->push_back('\\'); break; case '/': out->push_back('/'); break;
        case 'b': out->push_back('\b'); break; case 'f': out->push_back('\f'); break; case 'n': out->push_back('\n'); break;
        case 'r': out->push_back('\r'); break; case 't': out->push_back('\t'); break;
        case 'u': {
          uint32_t cp;
          if (!hex4(&cp)) return false;
          if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {   // surrogate pair
            p += 2;
            uint32_t lo;
            if (!hex4(&lo)) return false;
            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
          }
          append_utf8(*out, cp);
          break;
        }
        default: return fail("bad escape");
      }
    }
    if (p >= end) return fail("unterminated string");
    ++p;
    return true;
  }
  bool value(JVal* v, int depth = 0) {
    if (depth > 64) return fail("nesting too deep");
    ws();
    if (p >= end) return fail("unexpected end");
    if (*p == '{') {
      ++p; v->type = JVal::Obj; ws();
      if (p < end && *p == '}') { ++p; return true; }
      while (true) {
        ws();
        std::string k;
        if (!string(&k)) return false;
        ws();
        if (p >= end || *p != ':') return fail("expected ':'");
        ++p;
        v->o.emplace_back(std::move(k), JVal());
        if (!value(&v->o.back().second, depth + 1)) return false;
        ws();
        if (p < end && *p == ',') { ++p; continue; }
        if (p < end && *p == '}') { ++p; return true; }
        return fail("expected ',' or '}'");
      }
    }
    if (*p == '[') {
      ++p; v->type = JVal::Arr; ws();
      if (p < end && *p == ']') { ++p; return true; }
      while (true) {
        v->a.emplace_back();
        if (!value(&v->a.back(), depth + 1)) return false;
        ws();
        if (p < end && *p == ',') { ++p; continue; }
        if (p < end && *p == ']') { ++p; return true; }
        return fail("expected ',' or ']'");
      }
    }
    if (*p == '"') { v->type = JVal::Str; return string(&v->s); }
    if (!strncmp(p, "true", std::min<size_t>(4, end - p)) && end - p >= 4) { p += 4; v->type = JVal::Bool; v->b = true; return true; }
    if (!strncmp(p, "false", std::min<size_t>(5, end - p)) && end - p >= 5) { p += 5; v->type = JVal::Bool; v->b = false; return true; }
    if (!strncmp(p, "null", std::min<size_t>(4, end - p)) && end - p >= 4) { p += 4; v->type = JVal::Null; return true; }
    char* e2 = nullptr;
    v->n = strtod(p, &e2);
    if (e2 == p || e2 > end) return fail("bad number");
    p = e2;
    v->type = JVal::Num;
    return true;
  }
};

}  // namespace jsonmin
}  // namespace cl
