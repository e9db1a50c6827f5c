// See tflite_model.h.  Flatbuffer wire format: little-endian; a table starts with an int32 offset
// back to its vtable {u16 vtable_bytes, u16 table_bytes, u16 field_offset[...]}.  Field ids are those
// of the TFLite schema v3 (tensorflow/lite/schema/schema.fbs) and are named at each use.
#include "tflite_model.h"

#include <cstdio>
#include <fstream>

namespace lyra_b200 {
namespace {

class Cursor {
 public:
  Cursor(const std::vector<uint8_t>& img) : p_(img.data()), n_(img.size()) {}
  template <typename T> T Read(size_t off) const {
    if (off > n_ || sizeof(T) > n_ - off) throw std::runtime_error("tflite: read past end of file");
    T v; std::memcpy(&v, p_ + off, sizeof(T)); return v;
  }
  size_t Deref(size_t off) const { return off + Read<uint32_t>(off); }
  // absolute position of a table field, 0 if the field is absent (default value)
  size_t Field(size_t table, int id) const {
    const int64_t vts = (int64_t)table - (int64_t)Read<int32_t>(table);
    if (vts < 0 || (uint64_t)vts >= n_) throw std::runtime_error("tflite: vtable offset out of range");
    const size_t vt = (size_t)vts;
    const uint16_t vt_bytes = Read<uint16_t>(vt);
    const size_t slot = 4 + 2 * (size_t)id;
    if (slot + 2 > vt_bytes) return 0;
    const uint16_t rel = Read<uint16_t>(vt + slot);
    return rel ? table + rel : 0;
  }
  size_t SubTable(size_t table, int id) const { const size_t f = Field(table, id); return f ? Deref(f) : 0; }
  struct Vec { size_t begin = 0; uint32_t size = 0; };
  Vec Vector(size_t table, int id) const {
    const size_t f = Field(table, id);
    if (!f) return Vec{};
    const size_t v = Deref(f);
    return Vec{v + 4, Read<uint32_t>(v)};
  }
  template <typename T> std::vector<T> Scalars(size_t table, int id) const {
    const Vec v = Vector(table, id);
    if (v.begin > n_ || (uint64_t)v.size * sizeof(T) > n_ - v.begin) throw std::runtime_error("tflite: vector past end of file");
    std::vector<T> out(v.size);
    for (uint32_t i = 0; i < v.size; ++i) out[i] = Read<T>(v.begin + sizeof(T) * i);
    return out;
  }
  std::string String(size_t table, int id) const {
    const Vec v = Vector(table, id);
    if (v.begin > n_ || v.size > n_ - v.begin) throw std::runtime_error("tflite: string past end of file");
    return std::string(reinterpret_cast<const char*>(p_ + v.begin), v.size);
  }
  std::vector<size_t> Tables(size_t table, int id) const {
    const Vec v = Vector(table, id);
    if (v.begin > n_ || (uint64_t)v.size * 4 > n_ - v.begin) throw std::runtime_error("tflite: table vector past end of file");
    std::vector<size_t> out(v.size);
    for (uint32_t i = 0; i < v.size; ++i) out[i] = Deref(v.begin + 4 * (size_t)i);
    return out;
  }
  const uint8_t* base() const { return p_; }
  size_t size() const { return n_; }

 private:
  const uint8_t* p_;
  size_t n_;
};

}  // namespace

int TflSubgraph::producer(int tensor) const {
  for (size_t i = 0; i < ops.size(); ++i)
    for (int o : ops[i].outputs) if (o == tensor) return (int)i;
  return -1;
}
std::vector<int> TflSubgraph::consumers(int tensor) const {
  std::vector<int> out;
  for (size_t i = 0; i < ops.size(); ++i)
    for (int in : ops[i].inputs) if (in == tensor) { out.push_back((int)i); break; }
  return out;
}
int TflSubgraph::sole_consumer(int tensor, int code) const {
  int found = -1;
  for (int c : consumers(tensor))
    if (ops[c].code == code) { if (found >= 0) throw std::runtime_error("tflite: ambiguous consumer"); found = c; }
  if (found < 0) throw std::runtime_error("tflite: expected consumer op " + std::to_string(code) + " of tensor " + std::to_string(tensor));
  return found;
}

TflModel TflModel::Load(const std::string& path) {
  TflModel m;
  {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error("cannot open " + path);
    const std::streamsize n = f.tellg();
    if (n < 16) throw std::runtime_error("not a tflite file: " + path);
    m.image_.resize((size_t)n);
    f.seekg(0);
    f.read(reinterpret_cast<char*>(m.image_.data()), n);
  }
  if (std::memcmp(m.image_.data() + 4, "TFL3", 4) != 0) throw std::runtime_error("bad file identifier in " + path);
  const Cursor c(m.image_);
  const size_t model = c.Deref(0);

  // Model { 1: operator_codes, 2: subgraphs, 4: buffers, 7: signature_defs }
  std::vector<int> opcodes;
  for (size_t oc : c.Tables(model, 1)) {
    // OperatorCode { 0: deprecated_builtin_code:int8, 3: builtin_code:int32 } — the larger one is authoritative
    const size_t f0 = c.Field(oc, 0), f3 = c.Field(oc, 3);
    const int legacy = f0 ? c.Read<int8_t>(f0) : 0;
    const int modern = f3 ? c.Read<int32_t>(f3) : 0;
    opcodes.push_back(legacy > modern ? legacy : modern);
  }
  const std::vector<size_t> buffers = c.Tables(model, 4);

  for (size_t sg : c.Tables(model, 2)) {
    TflSubgraph g;
    // SubGraph { 0: tensors, 1: inputs, 2: outputs, 3: operators, 4: name }
    for (size_t tt : c.Tables(sg, 0)) {
      TflTensor t;
      // Tensor { 0: shape, 1: type, 2: buffer, 3: name, 4: quantization }
      t.shape = c.Scalars<int32_t>(tt, 0);
      const size_t ft = c.Field(tt, 1);
      t.type = (DType)(ft ? c.Read<int8_t>(ft) : 0);
      t.name = c.String(tt, 3);
      if (const size_t q = c.SubTable(tt, 4)) {
        // QuantizationParameters { 2: scale, 3: zero_point }
        t.scale = c.Scalars<float>(q, 2);
        t.zero_point = c.Scalars<int64_t>(q, 3);
      }
      const size_t fb = c.Field(tt, 2);
      const uint32_t bi = fb ? c.Read<uint32_t>(fb) : 0;
      if (bi < buffers.size()) {
        const Cursor::Vec d = c.Vector(buffers[bi], 0);   // Buffer { 0: data }
        if (d.size) {
          if (d.begin > c.size() || d.size > c.size() - d.begin) throw std::runtime_error("tflite: buffer past end of file");
          t.data = c.base() + d.begin;
          t.nbytes = d.size;
        }
      }
      g.tensors.push_back(std::move(t));
    }
    g.inputs = c.Scalars<int32_t>(sg, 1);
    g.outputs = c.Scalars<int32_t>(sg, 2);
    for (size_t ot : c.Tables(sg, 3)) {
      TflOp op;
      // Operator { 0: opcode_index, 1: inputs, 2: outputs, 4: builtin_options }
      const size_t fi = c.Field(ot, 0);
      const uint32_t idx = fi ? c.Read<uint32_t>(fi) : 0;
      if (idx >= opcodes.size()) throw std::runtime_error("tflite: bad opcode index");
      op.code = opcodes[idx];
      op.inputs = c.Scalars<int32_t>(ot, 1);
      op.outputs = c.Scalars<int32_t>(ot, 2);
      op.options = (uint32_t)c.SubTable(ot, 4);
      g.ops.push_back(std::move(op));
    }
    g.name = c.String(sg, 4);
    m.subgraphs_.push_back(std::move(g));
  }
  for (size_t sd : c.Tables(model, 7)) {
    // SignatureDef { 2: signature_key, 4: subgraph_index }
    const size_t fs = c.Field(sd, 4);
    m.signatures_.emplace_back(c.String(sd, 2), fs ? (int)c.Read<uint32_t>(fs) : 0);
  }
  return m;
}

int TflModel::SignatureSubgraph(const std::string& key) const {
  for (const auto& s : signatures_) if (s.first == key) return s.second;
  return -1;
}

int32_t TflModel::OptI32(const TflOp& op, int field, int32_t dflt) const {
  if (!op.options) return dflt;
  const Cursor c(image_);
  const size_t f = c.Field(op.options, field);
  return f ? c.Read<int32_t>(f) : dflt;
}
int TflModel::OptI8(const TflOp& op, int field, int dflt) const {
  if (!op.options) return dflt;
  const Cursor c(image_);
  const size_t f = c.Field(op.options, field);
  return f ? (int)c.Read<int8_t>(f) : dflt;
}
float TflModel::OptF32(const TflOp& op, int field, float dflt) const {
  if (!op.options) return dflt;
  const Cursor c(image_);
  const size_t f = c.Field(op.options, field);
  return f ? c.Read<float>(f) : dflt;
}
std::string TflModel::OptString(const TflOp& op, int field) const {
  if (!op.options) return std::string();
  const Cursor c(image_);
  return c.String(op.options, field);
}

}  // namespace lyra_b200
