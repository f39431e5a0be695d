// ms_json.h -- the JSON envelope of the Maelstrom protocol on the host side of the C ABI
// (ms_send_json / ms_recv_json): what process/parse-msg + keywordize-keys-1 + net/check-message do to
// every line a node prints (src/maelstrom/process.clj:26-66, net.clj:27-37) and what
// json/generate-stream does to every message a node reads (process.clj:162).
// A small recursive-descent parser that keeps the text of every value (so payloads the device does
// not interpret travel verbatim) and a writer with sorted body keys.
#pragma once
#include <stdint.h>
#include <stdlib.h>

#include <map>
#include <string>
#include <vector>

namespace msj {

struct Value {
  enum Kind { Null, Bool, Int, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  int64_t i = 0;
  std::string s;                                      // Str: decoded; Num: the literal
  std::vector<Value> arr;
  std::vector<std::pair<std::string, Value>> obj;     // in document order
  std::string text;                                   // the value's JSON text, verbatim

  const Value* get(const std::string& k) const {
    for (const auto& kv : obj) if (kv.first == k) return &kv.second;
    return nullptr;
  }
};

class Parser {
 public:
  explicit Parser(const std::string& src) : s_(src) {}
  bool parse(Value& out, std::string& err) {
    ws();
    if (!value(out, err)) return false;
    ws();
    if (p_ != s_.size()) { err = "trailing characters after the JSON value"; return false; }
    return true;
  }

 private:
  const std::string& s_;
  size_t p_ = 0;
  void ws() { while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\t' || s_[p_] == '\n' || s_[p_] == '\r')) p_++; }
  bool lit(const char* w) { const size_t n = strlen(w); if (s_.compare(p_, n, w) == 0) { p_ += n; return true; } return false; }
  static void utf8(std::string& o, uint32_t c) {
    if (c < 0x80) o += (char)c;
    else if (c < 0x800) { o += (char)(0xC0 | (c >> 6)); o += (char)(0x80 | (c & 0x3F)); }
    else if (c < 0x10000) { o += (char)(0xE0 | (c >> 12)); o += (char)(0x80 | ((c >> 6) & 0x3F)); o += (char)(0x80 | (c & 0x3F)); }
    else { o += (char)(0xF0 | (c >> 18)); o += (char)(0x80 | ((c >> 12) & 0x3F)); o += (char)(0x80 | ((c >> 6) & 0x3F)); o += (char)(0x80 | (c & 0x3F)); }
  }
  bool hex4(uint32_t& c) {
    if (p_ + 4 > s_.size()) return false;
    c = 0;
    for (int k = 0; k < 4; k++) {
      const char h = s_[p_++];
      c <<= 4;
      if (h >= '0' && h <= '9') c |= h - '0';
      else if (h >= 'a' && h <= 'f') c |= h - 'a' + 10;
      else if (h >= 'A' && h <= 'F') c |= h - 'A' + 10;
      else return false;
    }
    return true;
  }
  bool string(std::string& out, std::string& err) {
    p_++;   // opening quote
    while (p_ < s_.size()) {
      const char c = s_[p_++];
      if (c == '"') return true;
      if ((unsigned char)c < 0x20) break;
      if (c != '\\') { out += c; continue; }
      if (p_ >= s_.size()) break;
      const char e = s_[p_++];
      switch (e) {
        case '"': out += '"'; break; case '\\': out += '\\'; break; case '/': out += '/'; break;
        case 'b': out += '\b'; break; case 'f': out += '\f'; break; case 'n': out += '\n'; break;
        case 'r': out += '\r'; break; case 't': out += '\t'; break;
        case 'u': {
          uint32_t c1;
          if (!hex4(c1)) { err = "bad \\u escape"; return false; }
          if (c1 >= 0xD800 && c1 < 0xDC00 && p_ + 1 < s_.size() && s_[p_] == '\\' && s_[p_ + 1] == 'u') {
            p_ += 2;
            uint32_t c2;
            if (!hex4(c2)) { err = "bad \\u escape"; return false; }
            c1 = 0x10000 + ((c1 - 0xD800) << 10) + (c2 - 0xDC00);
          }
          utf8(out, c1);
          break;
        }
        default: err = "bad escape in string"; return false;
      }
    }
    err = "unterminated string";
    return false;
  }
  bool value(Value& v, std::string& err) {
    if (p_ >= s_.size()) { err = "unexpected end of input"; return false; }
    const size_t start = p_;
    const char c = s_[p_];
    bool ok = true;
    if (c == '{') {
      v.kind = Value::Obj;
      p_++; ws();
      if (p_ < s_.size() && s_[p_] == '}') p_++;
      else for (;;) {
        ws();
        if (p_ >= s_.size() || s_[p_] != '"') { err = "expected a string key"; return false; }
        std::string k;
        if (!string(k, err)) return false;
        ws();
        if (p_ >= s_.size() || s_[p_] != ':') { err = "expected ':'"; return false; }
        p_++; ws();
        Value child;
        if (!value(child, err)) return false;
        v.obj.emplace_back(k, child);
        ws();
        if (p_ < s_.size() && s_[p_] == ',') { p_++; continue; }
        if (p_ < s_.size() && s_[p_] == '}') { p_++; break; }
        err = "expected ',' or '}'";
        return false;
      }
    } else if (c == '[') {
      v.kind = Value::Arr;
      p_++; ws();
      if (p_ < s_.size() && s_[p_] == ']') p_++;
      else for (;;) {
        ws();
        Value child;
        if (!value(child, err)) return false;
        v.arr.push_back(child);
        ws();
        if (p_ < s_.size() && s_[p_] == ',') { p_++; continue; }
        if (p_ < s_.size() && s_[p_] == ']') { p_++; break; }
        err = "expected ',' or ']'";
        return false;
      }
    } else if (c == '"') {
      v.kind = Value::Str;
      ok = string(v.s, err);
    } else if (lit("true")) { v.kind = Value::Bool; v.b = true; }
    else if (lit("false")) { v.kind = Value::Bool; v.b = false; }
    else if (lit("null")) { v.kind = Value::Null; }
    else if (c == '-' || (c >= '0' && c <= '9')) {
      size_t q = p_;
      if (s_[q] == '-') q++;
      bool integral = true;
      while (q < s_.size() && ((s_[q] >= '0' && s_[q] <= '9') || s_[q] == '.' || s_[q] == 'e' || s_[q] == 'E' || s_[q] == '+' || s_[q] == '-')) {
        if (s_[q] == '.' || s_[q] == 'e' || s_[q] == 'E') integral = false;
        q++;
      }
      v.s = s_.substr(p_, q - p_);
      if (v.s == "-" || v.s.empty()) { err = "bad number"; return false; }
      v.kind = integral ? Value::Int : Value::Num;
      if (integral) v.i = strtoll(v.s.c_str(), nullptr, 10);
      p_ = q;
    } else {
      err = std::string("unexpected character '") + c + "'";
      return false;
    }
    if (!ok) return false;
    v.text = s_.substr(start, p_ - start);
    return true;
  }
};

inline std::string quote(const std::string& s) {
  std::string o = "\"";
  for (unsigned char c : s) {
    switch (c) {
      case '"': o += "\\\""; break; case '\\': o += "\\\\"; break; case '\n': o += "\\n"; break;
      case '\r': o += "\\r"; break; case '\t': o += "\\t"; break; case '\b': o += "\\b"; break; case '\f': o += "\\f"; break;
      default:
        if (c < 0x20) { char buf[8]; snprintf(buf, sizeof buf, "\\u%04x", c); o += buf; }
        else o += (char)c;
    }
  }
  return o + "\"";
}

// {"k":v,...} with keys in sorted order; values are JSON texts
inline std::string object(const std::map<std::string, std::string>& kv) {
  std::string o = "{";
  bool first = true;
  for (const auto& e : kv) {
    if (!first) o += ",";
    first = false;
    o += quote(e.first) + ":" + e.second;
  }
  return o + "}";
}

}  // namespace msj
