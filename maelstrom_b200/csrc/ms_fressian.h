// ms_fressian.h -- writes journal events in the reference's on-disk format: a stream of Fressian
// objects, one `Event{id time type message}` per event, exactly as maelstrom.net.journal's write
// handlers emit them (src/maelstrom/net/journal.clj:55-92):
//   "ev"  struct of 4: writeInt id, writeInt time, writeObject type (a keyword, cached), the message
//   "msg" struct of 4: writeInt id, writeObject src (cached), writeObject dest (cached), the body
//   body  tag "map" (code 0xC0) + closed list of key (keyword, cached) / value pairs; the value of
//         :type is cached too (write-body!, journal.clj:55-68)
// Wire format: org.fressian 0.6.x as used by clojure.data.fressian (a third-party dependency, not
// vendored under /root/reference; restated from its published encoding: packed ints, packed-length
// strings, STRUCTTYPE / struct cache, PUT_PRIORITY_CACHE / priority cache, "key" = code 0xCA with
// two cached components (namespace, name)).  No JVM exists in the build image: the encoder is pinned
// by a hand-decoded byte fixture and a reader written from the same description
// (tests/test_fressian_journal.py), not by the reference's reader -- "parity unpinned" for this file.
#pragma once
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <unordered_map>
#include <vector>

namespace msf {

enum : uint8_t {
  C_PRIORITY_CACHE_PACKED_START = 0x80, C_STRUCT_CACHE_PACKED_START = 0xA0, C_MAP = 0xC0, C_KEY = 0xCA,
  C_GET_PRIORITY_CACHE = 0xCC, C_PUT_PRIORITY_CACHE = 0xCD, C_STRING_PACKED_LENGTH_START = 0xDA, C_STRING = 0xE3,
  C_BEGIN_CLOSED_LIST = 0xED, C_STRUCTTYPE = 0xEF, C_STRUCT = 0xF0, C_NULL = 0xF7, C_INT = 0xF8, C_END_COLLECTION = 0xFD
};

class Writer {
 public:
  explicit Writer(FILE* f) : f_(f) {}

  // FressianWriter.writeInt: shortest of the packed forms by the number of significant bits
  void write_int(int64_t i) {
    const uint64_t m = (uint64_t)(i < 0 ? ~i : i);
    const int lz = m ? __builtin_clzll(m) : 64;
    if (lz <= 14) { raw(C_INT); raw_be(i, 8); }
    else if (lz <= 22) { raw((uint8_t)(0x7E + (i >> 48))); raw_be(i, 6); }
    else if (lz <= 30) { raw((uint8_t)(0x7A + (i >> 40))); raw_be(i, 5); }
    else if (lz <= 38) { raw((uint8_t)(0x76 + (i >> 32))); raw_be(i, 4); }
    else if (lz <= 44) { raw((uint8_t)(0x72 + (i >> 24))); raw_be(i, 3); }
    else if (lz <= 51) { raw((uint8_t)(0x68 + (i >> 16))); raw_be(i, 2); }
    else if (lz <= 57 || i < -1) { raw((uint8_t)(0x50 + (i >> 8))); raw_be(i, 1); }
    else raw((uint8_t)i);                                   // -1 .. 63 in one byte
  }

  // FressianWriter.writeString (ASCII / UTF-8 bytes as they are; node names and type names are ASCII)
  void write_string(const std::string& s) {
    if (s.size() < 8) raw((uint8_t)(C_STRING_PACKED_LENGTH_START + s.size()));
    else { raw(C_STRING); write_int((int64_t)s.size()); }
    fwrite(s.data(), 1, s.size(), f_);
  }

  // writeObject(o, true) for a String
  void write_string_cached(const std::string& s) {
    if (s.empty()) { write_string(s); return; }             // shouldSkipCache
    if (cached("s:" + s)) return;
    raw(C_PUT_PRIORITY_CACHE);
    write_string(s);
  }

  // writeObject(kw, true) for an un-namespaced Clojure keyword: tag "key" (code 0xCA), components
  // (namespace = nil, name), both written through the cache as well
  void write_keyword_cached(const std::string& name) {
    if (cached("k:" + name)) return;
    raw(C_PUT_PRIORITY_CACHE);
    raw(C_KEY);
    raw(C_NULL);
    write_string_cached(name);
  }

  // writeTag for a tag without a built-in code
  void write_struct_tag(const std::string& tag, int components) {
    auto it = structs_.find(tag);
    if (it == structs_.end()) {
      structs_.emplace(tag, (int)structs_.size());
      raw(C_STRUCTTYPE);
      write_string(tag);
      write_int(components);
    } else if (it->second < 16) {
      raw((uint8_t)(C_STRUCT_CACHE_PACKED_START + it->second));
    } else {
      raw(C_STRUCT);
      write_int(it->second);
    }
  }

  struct KV { const char* key; bool is_string; int64_t i; std::string s; };

  // one journal event (journal.clj:70-92)
  void write_event(int64_t id, int64_t time_ns, bool recv, int64_t msg_id, const std::string& src,
                   const std::string& dest, const std::vector<KV>& body) {
    write_struct_tag("ev", 4);
    write_int(id);
    write_int(time_ns);
    write_keyword_cached(recv ? "recv" : "send");
    write_struct_tag("msg", 4);
    write_int(msg_id);
    write_string_cached(src);
    write_string_cached(dest);
    raw(C_MAP);                                              // (.writeTag w "map" 1)
    raw(C_BEGIN_CLOSED_LIST);
    for (const KV& kv : body) {
      write_keyword_cached(kv.key);
      if (!kv.is_string) write_int(kv.i);
      else if (std::string(kv.key) == "type") write_string_cached(kv.s);
      else write_string(kv.s);
    }
    raw(C_END_COLLECTION);
  }

 private:
  FILE* f_;
  std::unordered_map<std::string, int> cache_, structs_;

  void raw(uint8_t b) { fputc(b, f_); }
  void raw_be(int64_t v, int bytes) {
    for (int k = bytes - 1; k >= 0; k--) fputc((int)(((uint64_t)v >> (8 * k)) & 0xFF), f_);
  }
  // writes the reference to an already cached object and returns true, or interns it and returns false
  bool cached(const std::string& key) {
    auto it = cache_.find(key);
    if (it == cache_.end()) { cache_.emplace(key, (int)cache_.size()); return false; }
    if (it->second < 32) raw((uint8_t)(C_PRIORITY_CACHE_PACKED_START + it->second));
    else { raw(C_GET_PRIORITY_CACHE); write_int(it->second); }
    return true;
  }
};

}  // namespace msf
