// filter.cu -- device-side evaluation of the reference's MetadataFilter (src/storage.rs:44-71).
//
// The reference keeps `HashMap<usize, Metadata{HashMap<String,String>}>` in VectorStore and walks it on
// the host for every candidate (storage.rs:272-285).  Here every metadata field is a dictionary-encoded
// u32 column in HBM (code 0 = field absent); a filter is compiled to a short postfix program and one
// kernel turns it into the eligibility bitmask (one bit per slot) that K1/K2 consume -- so a filtered
// search over 10M rows never touches host metadata.  Truth table = storage.rs:62-70:
//   Eq: get(field) == Some(value) | Ne: get(field) != Some(value) (true when absent) | Exists |
//   And = all (empty => true) | Or = any (empty => false).
#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace gfi {

namespace {

__global__ void eval_filter_kernel(const FilterProgram prog, const uint32_t* const* cols, int64_t cols_len,
                                   int64_t n_slots, uint64_t* mask_words) {
  griddep_wait();
  // one thread per slot; the 64 slots of a mask word are assembled with two ballots
  const int lane = threadIdx.x & 31;
  const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ll;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t s0 = base; s0 < ((n_slots + 31) & ~31ll); s0 += stride) {
    const int64_t s = s0 + lane;
    bool stack[kFilterMaxDepth];
    int sp = 0;
    bool res = false;
    if (s < n_slots) {
      for (int i = 0; i < prog.n; ++i) {
        const FilterOp op = prog.ops[i];
        if (op.kind <= kFilterExists) {
          const uint32_t code = (op.field >= 0 && s < cols_len) ? cols[op.field][s] : 0u;  // beyond the synced columns: absent
          bool v;
          if (op.kind == kFilterEq) v = code != 0u && code == op.code;
          else if (op.kind == kFilterNe) v = !(code != 0u && code == op.code);
          else v = code != 0u;
          stack[sp++] = v;
        } else {
          bool acc = op.kind == kFilterAnd;
          for (uint32_t c = 0; c < op.code; ++c) {
            const bool v = stack[--sp];
            acc = op.kind == kFilterAnd ? (acc && v) : (acc || v);
          }
          stack[sp++] = acc;
        }
      }
      res = sp > 0 ? stack[sp - 1] : true;
    }
    const unsigned m = __ballot_sync(0xffffffffu, res);
    if (lane == 0) reinterpret_cast<uint32_t*>(mask_words)[s0 >> 5] = m;
  }
}

// ---- a minimal JSON reader for the serde form {"op": "...", "field": "...", "value": "...", "filters": [...]} ----
struct Json {
  const char* p;
  const char* e;
  std::string err;
  void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
  bool lit(char c) { ws(); if (p < e && *p == c) { ++p; return true; } return false; }
  bool str(std::string& out) {
    ws();
    if (p >= e || *p != '"') { err = "expected string"; return false; }
    ++p;
    out.clear();
    while (p < e && *p != '"') {
      if (*p == '\\' && p + 1 < e) {
        ++p;
        switch (*p) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            if (p + 4 >= e) { err = "bad \\u escape"; return false; }
            unsigned v = 0;
            for (int i = 1; i <= 4; ++i) {
              const char c = p[i];
              v = v * 16 + (c >= '0' && c <= '9' ? c - '0' : (c | 32) - 'a' + 10);
            }
            p += 4;
            if (v < 0x80) out += (char)v;
            else if (v < 0x800) { out += (char)(0xC0 | (v >> 6)); out += (char)(0x80 | (v & 0x3F)); }
            else { out += (char)(0xE0 | (v >> 12)); out += (char)(0x80 | ((v >> 6) & 0x3F)); out += (char)(0x80 | (v & 0x3F)); }
            break;
          }
          default: out += *p;
        }
        ++p;
      } else {
        out += *p++;
      }
    }
    if (p >= e) { err = "unterminated string"; return false; }
    ++p;
    return true;
  }
};

}  // namespace

struct FilterCompiler {
  const std::map<std::string, int>& fields;
  const std::vector<std::map<std::string, uint32_t>>& values;
  FilterProgram prog{};
  std::string err;

  bool emit(int kind, int field, uint32_t code) {
    if (prog.n >= kFilterMaxOps) { err = "filter too large"; return false; }
    prog.ops[prog.n++] = FilterOp{kind, field, code};
    return true;
  }
  // parses one filter object, emits postfix code; depth = operand-stack depth before this node
  bool node(Json& j, int depth) {
    if (depth >= kFilterMaxDepth) { err = "filter nested too deeply"; return false; }
    if (!j.lit('{')) { err = "expected '{'"; return false; }
    std::string op, field, value;
    bool have_filters = false;
    uint32_t nchildren = 0;
    bool first = true;
    while (!j.lit('}')) {
      if (!first && !j.lit(',')) { err = "expected ','"; return false; }
      first = false;
      std::string key;
      if (!j.str(key)) { err = j.err; return false; }
      if (!j.lit(':')) { err = "expected ':'"; return false; }
      if (key == "op") { if (!j.str(op)) { err = j.err; return false; } }
      else if (key == "field") { if (!j.str(field)) { err = j.err; return false; } }
      else if (key == "value") { if (!j.str(value)) { err = j.err; return false; } }
      else if (key == "filters") {
        have_filters = true;
        if (!j.lit('[')) { err = "expected '['"; return false; }
        if (!j.lit(']')) {
          do {
            // children are evaluated left to right, each leaving one value on the stack
            if (!node(j, depth + (int)nchildren)) return false;
            ++nchildren;
          } while (j.lit(','));
          if (!j.lit(']')) { err = "expected ']'"; return false; }
        }
      } else { err = "unknown key: " + key; return false; }
    }
    if (op == "eq" || op == "ne" || op == "exists") {
      int f = -1;
      uint32_t code = 0xffffffffu;  // a value never inserted matches no row
      auto it = fields.find(field);
      if (it != fields.end()) {
        f = it->second;
        auto vt = values[f].find(value);
        if (vt != values[f].end()) code = vt->second;
      }
      return emit(op == "eq" ? kFilterEq : op == "ne" ? kFilterNe : kFilterExists, f, code);
    }
    if (op == "and" || op == "or") {
      if (!have_filters) { err = "missing 'filters'"; return false; }
      return emit(op == "and" ? kFilterAnd : kFilterOr, -1, nchildren);
    }
    err = "unknown op: " + op;
    return false;
  }
};

bool compile_filter(const char* json, const std::map<std::string, int>& fields,
                    const std::vector<std::map<std::string, uint32_t>>& values, FilterProgram* out, std::string* err) {
  Json j{json, json + strlen(json), {}};
  FilterCompiler c{fields, values};
  if (!c.node(j, 0)) { *err = c.err.empty() ? j.err : c.err; return false; }
  j.ws();
  if (j.p != j.e) { *err = "trailing characters after filter"; return false; }
  *out = c.prog;
  return true;
}

cudaError_t launch_eval_filter(const FilterProgram& prog, const uint32_t* const* d_cols, int64_t cols_len,
                               int64_t n_slots, uint64_t* mask_words, cudaStream_t st) {
  if (n_slots <= 0) return cudaSuccess;
  const int blocks = (int)std::min<int64_t>((n_slots + 255) / 256, 148 * 8);
  return launch_pdl(eval_filter_kernel, dim3(blocks), dim3(256), 0, st, prog, d_cols, cols_len, n_slots, mask_words);
}

}  // namespace gfi
