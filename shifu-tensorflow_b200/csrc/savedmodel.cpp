// SavedModel (saved_model.pb) + tensor-bundle V2 (variables/variables.{index,data-00000-of-00001})
// writer and reader, host only, no TensorFlow / protobuf dependency.
//
// Writer  = simple_save + export_generic_config of the reference trainer
//           (shifu-tensorflow-on-yarn/src/main/resources/ssgd_monitor.py:457-490): tag "serve", signature
//           "serving_default" (predict: shifu_input_0 -> shifu_output_0), node / variable names of
//           nn_layer (:57-71): weight_<name>, biases_<name>, MatMul[_k], add[_k], hidden_layer<k>, shifu_output_0.
// Reader  = what SavedModelBundle.load + feed/fetch by op name need
//           (shifu-tensorflow-eval/src/main/java/ml/shifu/shifu/tensorflow/TensorflowModel.java:71,85,169):
//           walk the graph from the fetched op back to the fed placeholder through
//           activation <- BiasAdd|Add <- MatMul chains and pull the kernels out of the bundle.  Handles the
//           Keras-exported fixture (dropout gated by a learning-phase Switch/Merge) as well.
//
// Formats: protobuf wire format; leveldb table (prefix-compressed blocks with restart arrays, 5-byte block
// trailer {compression type, masked crc32c}, 48-byte footer, magic 0xdb4775248b80fb57); TF
// tensor_bundle.proto (BundleHeaderProto, BundleEntryProto).
#include <errno.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "savedmodel.h"

namespace sb {
int set_error(int code, const char* fmt, ...);
}
using sb::set_error;

namespace {

typedef std::string Bytes;

// ------------------------------------------------------------------ protobuf encode
void put_varint(Bytes& b, uint64_t v) {
  while (v >= 0x80) { b.push_back(static_cast<char>((v & 0x7F) | 0x80)); v >>= 7; }
  b.push_back(static_cast<char>(v));
}
void put_tag(Bytes& b, int field, int wt) { put_varint(b, (static_cast<uint64_t>(field) << 3) | wt); }
void put_int(Bytes& b, int field, int64_t v) { put_tag(b, field, 0); put_varint(b, static_cast<uint64_t>(v)); }
void put_bytes(Bytes& b, int field, const Bytes& s) { put_tag(b, field, 2); put_varint(b, s.size()); b += s; }
void put_fixed32(Bytes& b, int field, uint32_t v) { put_tag(b, field, 5); b.append(reinterpret_cast<const char*>(&v), 4); }
void put_float(Bytes& b, int field, float f) { uint32_t u; memcpy(&u, &f, 4); put_fixed32(b, field, u); }

enum { DT_FLOAT = 1, DT_INT32 = 3, DT_STRING = 7, DT_BOOL = 10 };

Bytes shape_proto(const std::vector<int64_t>& dims) {
  Bytes s;
  for (int64_t d : dims) { Bytes dm; put_int(dm, 1, d); put_bytes(s, 2, dm); }
  return s;
}
Bytes attr_type(int dt) { Bytes a; put_int(a, 6, dt); return a; }
Bytes attr_shape(const std::vector<int64_t>& dims) { Bytes a; put_bytes(a, 7, shape_proto(dims)); return a; }
Bytes attr_bool(bool v) { Bytes a; put_int(a, 5, v ? 1 : 0); return a; }
Bytes attr_str(const Bytes& s) { Bytes a; put_bytes(a, 2, s); return a; }
Bytes attr_f(float f) { Bytes a; put_float(a, 4, f); return a; }
Bytes attr_str_list(const std::vector<Bytes>& v) { Bytes l; for (auto& s : v) put_bytes(l, 2, s); Bytes a; put_bytes(a, 1, l); return a; }
Bytes attr_type_list(int dt, int n) {
  Bytes packed; for (int i = 0; i < n; ++i) put_varint(packed, dt);
  Bytes l; put_bytes(l, 6, packed);
  Bytes a; put_bytes(a, 1, l); return a;
}
Bytes attr_string_tensor(const std::vector<Bytes>& vals, bool scalar) {
  Bytes t; put_int(t, 1, DT_STRING);
  put_bytes(t, 2, scalar ? Bytes() : shape_proto({static_cast<int64_t>(vals.size())}));
  for (auto& v : vals) put_bytes(t, 8, v);
  Bytes a; put_bytes(a, 8, t); return a;
}

struct NodeW {
  Bytes name, op;
  std::vector<Bytes> inputs;
  std::vector<std::pair<Bytes, Bytes>> attrs;  // name -> serialized AttrValue
};
Bytes node_proto(const NodeW& n) {
  Bytes b;
  put_bytes(b, 1, n.name);
  put_bytes(b, 2, n.op);
  for (auto& i : n.inputs) put_bytes(b, 3, i);
  auto attrs = n.attrs;
  std::sort(attrs.begin(), attrs.end());
  for (auto& kv : attrs) { Bytes e; put_bytes(e, 1, kv.first); put_bytes(e, 2, kv.second); put_bytes(b, 5, e); }
  return b;
}

// ------------------------------------------------------------------ crc32c
uint32_t crc_table[256];
bool crc_init = false;
uint32_t crc32c(const void* data, size_t n, uint32_t crc = 0) {
  if (!crc_init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      crc_table[i] = c;
    }
    crc_init = true;
  }
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = crc ^ 0xFFFFFFFFu;
  for (size_t i = 0; i < n; ++i) c = crc_table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}
uint32_t crc_mask(uint32_t c) { return ((c >> 15) | (c << 17)) + 0xA282EAD8u; }

// ------------------------------------------------------------------ leveldb table writer
const uint64_t TABLE_MAGIC = 0xDB4775248B80FB57ull;

struct BlockBuilder {
  Bytes buf, last_key;
  std::vector<uint32_t> restarts{0};
  int counter = 0;
  void add(const Bytes& key, const Bytes& val) {
    size_t shared = 0;
    if (counter < 16) {
      const size_t m = std::min(last_key.size(), key.size());
      while (shared < m && last_key[shared] == key[shared]) ++shared;
    } else {
      restarts.push_back(static_cast<uint32_t>(buf.size()));
      counter = 0;
    }
    put_varint(buf, shared);
    put_varint(buf, key.size() - shared);
    put_varint(buf, val.size());
    buf.append(key, shared, Bytes::npos);
    buf += val;
    last_key = key;
    ++counter;
  }
  Bytes finish() {
    Bytes b = buf;
    for (uint32_t r : restarts) b.append(reinterpret_cast<const char*>(&r), 4);
    uint32_t n = static_cast<uint32_t>(restarts.size());
    b.append(reinterpret_cast<const char*>(&n), 4);
    return b;
  }
};
// appends block + trailer to file image, returns handle (offset, size)
std::pair<uint64_t, uint64_t> emit_block(Bytes& file, const Bytes& block) {
  const uint64_t off = file.size();
  file += block;
  file.push_back(0);  // kNoCompression
  uint32_t c = crc_mask(crc32c(file.data() + off, block.size() + 1));
  file.append(reinterpret_cast<const char*>(&c), 4);
  return {off, block.size()};
}
Bytes handle_bytes(std::pair<uint64_t, uint64_t> h) { Bytes b; put_varint(b, h.first); put_varint(b, h.second); return b; }

Bytes build_table(const std::vector<std::pair<Bytes, Bytes>>& sorted_entries) {
  Bytes file;
  BlockBuilder data;
  for (auto& kv : sorted_entries) data.add(kv.first, kv.second);
  auto dh = emit_block(file, data.finish());
  BlockBuilder meta;
  auto mh = emit_block(file, meta.finish());
  BlockBuilder index;
  index.add(sorted_entries.empty() ? Bytes() : sorted_entries.back().first, handle_bytes(dh));
  auto ih = emit_block(file, index.finish());
  Bytes footer = handle_bytes(mh) + handle_bytes(ih);
  footer.resize(40, 0);
  footer.append(reinterpret_cast<const char*>(&TABLE_MAGIC), 8);
  file += footer;
  return file;
}

// ------------------------------------------------------------------ file helpers
int write_file(const std::string& path, const Bytes& data) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return set_error(SB_ERR_IO, "cannot open %s for writing: %s", path.c_str(), strerror(errno));
  bool ok = data.empty() || fwrite(data.data(), 1, data.size(), f) == data.size();
  ok = (fclose(f) == 0) && ok;
  if (!ok) return set_error(SB_ERR_IO, "short write to %s", path.c_str());
  return SB_OK;
}
int read_file(const std::string& path, Bytes* out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return set_error(SB_ERR_IO, "cannot open %s: %s", path.c_str(), strerror(errno));
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out->resize(static_cast<size_t>(n));
  bool ok = n == 0 || fread(&(*out)[0], 1, static_cast<size_t>(n), f) == static_cast<size_t>(n);
  fclose(f);
  if (!ok) return set_error(SB_ERR_IO, "short read from %s", path.c_str());
  return SB_OK;
}
int mkdir_p(const std::string& path) {
  std::string cur;
  for (size_t i = 0; i <= path.size(); ++i) {
    if (i == path.size() || path[i] == '/') {
      if (!cur.empty() && mkdir(cur.c_str(), 0755) != 0 && errno != EEXIST)
        return set_error(SB_ERR_IO, "mkdir %s failed: %s", cur.c_str(), strerror(errno));
    }
    if (i < path.size()) cur.push_back(path[i]);
  }
  return SB_OK;
}

const char* act_op_name(int act) {
  switch (act) {
    case SB_ACT_SIGMOID: return "Sigmoid";
    case SB_ACT_TANH: return "Tanh";
    case SB_ACT_RELU: return "Relu";
    case SB_ACT_LEAKYRELU: return "LeakyRelu";
    default: return "Identity";
  }
}

// ------------------------------------------------------------------ protobuf decode
struct Field { int no; int wt; uint64_t ival; const char* p; size_t len; };
bool get_varint(const char*& p, const char* end, uint64_t* v) {
  uint64_t r = 0; int s = 0;
  while (p < end) {
    uint8_t b = static_cast<uint8_t>(*p++);
    r |= static_cast<uint64_t>(b & 0x7F) << s;
    if (!(b & 0x80)) { *v = r; return true; }
    s += 7;
    if (s > 63) return false;
  }
  return false;
}
bool parse_msg(const char* p, size_t n, std::vector<Field>* out) {
  const char* end = p + n;
  while (p < end) {
    uint64_t key;
    if (!get_varint(p, end, &key)) return false;
    Field f = {static_cast<int>(key >> 3), static_cast<int>(key & 7), 0, nullptr, 0};
    if (f.wt == 0) { if (!get_varint(p, end, &f.ival)) return false; }
    else if (f.wt == 1) { if (end - p < 8) return false; memcpy(&f.ival, p, 8); p += 8; }
    else if (f.wt == 5) { if (end - p < 4) return false; uint32_t u; memcpy(&u, p, 4); f.ival = u; p += 4; }
    else if (f.wt == 2) {
      uint64_t len;
      if (!get_varint(p, end, &len) || static_cast<uint64_t>(end - p) < len) return false;
      f.p = p; f.len = static_cast<size_t>(len); p += len;
    } else return false;
    out->push_back(f);
  }
  return true;
}
std::string fstr(const Field& f) { return std::string(f.p, f.len); }

struct NodeR { std::string op; std::vector<std::string> inputs; };

std::string strip_name(const std::string& s) {
  size_t b = (!s.empty() && s[0] == '^') ? 1 : 0;
  size_t c = s.find(':', b);
  return s.substr(b, c == std::string::npos ? std::string::npos : c - b);
}

struct BundleEntry { int dtype = 0; std::vector<int64_t> shape; int64_t offset = 0, size = 0; };

int read_table_entries(const Bytes& file, std::vector<std::pair<Bytes, Bytes>>* out) {
  if (file.size() < 48) return set_error(SB_ERR_FORMAT, "bundle index too small");
  uint64_t magic;
  memcpy(&magic, file.data() + file.size() - 8, 8);
  if (magic != TABLE_MAGIC) return set_error(SB_ERR_FORMAT, "bundle index: bad table magic");
  const char* p = file.data() + file.size() - 48;
  const char* end = p + 40;
  uint64_t mo, ms, io, is;
  if (!get_varint(p, end, &mo) || !get_varint(p, end, &ms) || !get_varint(p, end, &io) || !get_varint(p, end, &is))
    return set_error(SB_ERR_FORMAT, "bundle index: bad footer");
  auto read_block = [&](uint64_t off, uint64_t size, std::vector<std::pair<Bytes, Bytes>>* ents) -> int {
    if (off + size + 5 > file.size()) return set_error(SB_ERR_FORMAT, "bundle index: block out of range");
    if (file[off + size] != 0) return set_error(SB_ERR_FORMAT, "bundle index: compressed blocks unsupported");
    uint32_t stored;
    memcpy(&stored, file.data() + off + size + 1, 4);
    if (crc_mask(crc32c(file.data() + off, size + 1)) != stored) return set_error(SB_ERR_FORMAT, "bundle index: block crc mismatch");
    if (size < 4) return set_error(SB_ERR_FORMAT, "bundle index: short block");
    uint32_t nr;
    memcpy(&nr, file.data() + off + size - 4, 4);
    if (static_cast<uint64_t>(nr) * 4 + 4 > size) return set_error(SB_ERR_FORMAT, "bundle index: bad restart count");
    const char* q = file.data() + off;
    const char* qe = q + size - 4 - 4ull * nr;
    Bytes key;
    while (q < qe) {
      uint64_t sh, ns, vl;
      if (!get_varint(q, qe, &sh) || !get_varint(q, qe, &ns) || !get_varint(q, qe, &vl) || sh > key.size() ||
          static_cast<uint64_t>(qe - q) < ns + vl)
        return set_error(SB_ERR_FORMAT, "bundle index: corrupt entry");
      key = key.substr(0, sh) + Bytes(q, ns);
      q += ns;
      ents->push_back({key, Bytes(q, vl)});
      q += vl;
    }
    return SB_OK;
  };
  std::vector<std::pair<Bytes, Bytes>> idx;
  int s = read_block(io, is, &idx);
  if (s != SB_OK) return s;
  for (auto& kv : idx) {
    const char* hp = kv.second.data();
    const char* he = hp + kv.second.size();
    uint64_t bo, bs;
    if (!get_varint(hp, he, &bo) || !get_varint(hp, he, &bs)) return set_error(SB_ERR_FORMAT, "bundle index: bad handle");
    s = read_block(bo, bs, out);
    if (s != SB_OK) return s;
  }
  return SB_OK;
}

}  // namespace

extern "C" {

int sb_savedmodel_write(const char* export_dir, const sb_net_desc* d, const float* flat, int64_t n) {
  if (!export_dir || !d || !flat) return set_error(SB_ERR_INVALID, "null argument");
  if (d->n_features <= 0 || d->n_hidden < 1 || d->n_hidden > SB_MAX_HIDDEN) return set_error(SB_ERR_INVALID, "bad topology");
  const int L = d->n_hidden;
  // ---- variables in graph order ----
  struct Var { std::string name; std::vector<int64_t> shape; const float* data; int64_t count; };
  std::vector<Var> vars;
  {
    int64_t off = 0;
    int prev = d->n_features;
    for (int l = 0; l <= L; ++l) {
      const int out = l < L ? d->hidden[l] : 1;
      const std::string nm = l < L ? "hidden_layer" + std::to_string(l) : "shifu_output_0";
      vars.push_back({"weight_" + nm, {prev, out}, flat + off, static_cast<int64_t>(prev) * out}); off += static_cast<int64_t>(prev) * out;
      vars.push_back({"biases_" + nm, {out}, flat + off, out}); off += out;
      prev = out;
    }
    if (off != n) return set_error(SB_ERR_INVALID, "expected %lld params, got %lld", (long long)off, (long long)n);
  }
  // ---- graph ----
  std::vector<NodeW> nodes;
  nodes.push_back({"shifu_input_0", "Placeholder", {}, {{"dtype", attr_type(DT_FLOAT)}, {"shape", attr_shape({-1, d->n_features})}}});
  for (auto& v : vars) {
    nodes.push_back({v.name, "VariableV2", {}, {{"dtype", attr_type(DT_FLOAT)}, {"shape", attr_shape(v.shape)},
                                                 {"container", attr_str("")}, {"shared_name", attr_str("")}}});
    nodes.push_back({v.name + "/read", "Identity", {v.name}, {{"T", attr_type(DT_FLOAT)}, {"_class", attr_str_list({"loc:@" + v.name})}}});
  }
  std::string prev_node = "shifu_input_0";
  for (int l = 0; l <= L; ++l) {
    const std::string sfx = l == 0 ? "" : "_" + std::to_string(l);
    const std::string nm = l < L ? "hidden_layer" + std::to_string(l) : "shifu_output_0";
    const int act = l < L ? d->acts[l] : SB_ACT_SIGMOID;
    nodes.push_back({"MatMul" + sfx, "MatMul", {prev_node, "weight_" + nm + "/read"},
                     {{"T", attr_type(DT_FLOAT)}, {"transpose_a", attr_bool(false)}, {"transpose_b", attr_bool(false)}}});
    nodes.push_back({"add" + sfx, "Add", {"MatMul" + sfx, "biases_" + nm + "/read"}, {{"T", attr_type(DT_FLOAT)}}});
    NodeW a = {nm, act_op_name(act), {"add" + sfx}, {{"T", attr_type(DT_FLOAT)}}};
    if (act == SB_ACT_LEAKYRELU) a.attrs.push_back({"alpha", attr_f(0.2f)});
    nodes.push_back(a);
    prev_node = nm;
  }
  // restore sub-graph (what SavedModelBundle.load runs: restore_op_name with filename_tensor fed)
  std::vector<Bytes> vnames, empties;
  for (auto& v : vars) { vnames.push_back(v.name); empties.push_back(""); }
  nodes.push_back({"save/Const", "Const", {}, {{"dtype", attr_type(DT_STRING)}, {"value", attr_string_tensor({"model"}, true)}}});
  nodes.push_back({"save/RestoreV2/tensor_names", "Const", {}, {{"dtype", attr_type(DT_STRING)}, {"value", attr_string_tensor(vnames, false)}}});
  nodes.push_back({"save/RestoreV2/shape_and_slices", "Const", {}, {{"dtype", attr_type(DT_STRING)}, {"value", attr_string_tensor(empties, false)}}});
  nodes.push_back({"save/RestoreV2", "RestoreV2", {"save/Const", "save/RestoreV2/tensor_names", "save/RestoreV2/shape_and_slices"},
                   {{"dtypes", attr_type_list(DT_FLOAT, static_cast<int>(vars.size()))}}});
  NodeW restore_all = {"save/restore_all", "NoOp", {}, {}};
  for (size_t i = 0; i < vars.size(); ++i) {
    const std::string an = i == 0 ? "save/Assign" : "save/Assign_" + std::to_string(i);
    const std::string src = i == 0 ? "save/RestoreV2" : "save/RestoreV2:" + std::to_string(i);
    nodes.push_back({an, "Assign", {vars[i].name, src},
                     {{"T", attr_type(DT_FLOAT)}, {"_class", attr_str_list({"loc:@" + vars[i].name})},
                      {"use_locking", attr_bool(true)}, {"validate_shape", attr_bool(true)}}});
    restore_all.inputs.push_back("^" + an);
  }
  nodes.push_back(restore_all);

  Bytes graph;
  for (auto& nd : nodes) put_bytes(graph, 1, node_proto(nd));
  { Bytes ver; put_int(ver, 1, 24); put_bytes(graph, 4, ver); }

  Bytes meta_info;
  put_bytes(meta_info, 4, "serve");
  put_bytes(meta_info, 5, "1.4.0");
  put_bytes(meta_info, 6, "shifu_b200");

  Bytes saver;
  put_bytes(saver, 1, "save/Const:0");
  put_bytes(saver, 2, "save/Const:0");
  put_bytes(saver, 3, "save/restore_all");
  put_int(saver, 4, 5);
  put_float(saver, 6, 10000.0f);
  put_int(saver, 7, 2);

  auto tensor_info = [&](const std::string& name, int64_t cols) {
    Bytes t; put_bytes(t, 1, name); put_int(t, 2, DT_FLOAT); put_bytes(t, 3, shape_proto({-1, cols})); return t;
  };
  Bytes sig;
  { Bytes e; put_bytes(e, 1, "shifu_input_0"); put_bytes(e, 2, tensor_info("shifu_input_0:0", d->n_features)); put_bytes(sig, 1, e); }
  { Bytes e; put_bytes(e, 1, "shifu_output_0"); put_bytes(e, 2, tensor_info("shifu_output_0:0", 1)); put_bytes(sig, 2, e); }
  put_bytes(sig, 3, "tensorflow/serving/predict");

  Bytes mg;
  put_bytes(mg, 1, meta_info);
  put_bytes(mg, 2, graph);
  put_bytes(mg, 3, saver);
  { Bytes e; put_bytes(e, 1, "serving_default"); put_bytes(e, 2, sig); put_bytes(mg, 5, e); }

  Bytes sm;
  put_int(sm, 1, 1);
  put_bytes(sm, 2, mg);

  // ---- tensor bundle ----
  std::vector<Var> sorted = vars;
  std::sort(sorted.begin(), sorted.end(), [](const Var& a, const Var& b) { return a.name < b.name; });
  Bytes data;
  std::vector<std::pair<Bytes, Bytes>> entries;
  {
    Bytes hdr; put_int(hdr, 1, 1);
    Bytes ver; put_int(ver, 1, 1); put_bytes(hdr, 3, ver);
    entries.push_back({"", hdr});
  }
  for (auto& v : sorted) {
    const int64_t off = static_cast<int64_t>(data.size());
    const int64_t size = v.count * 4;
    data.append(reinterpret_cast<const char*>(v.data), static_cast<size_t>(size));
    Bytes e;
    put_int(e, 1, DT_FLOAT);
    put_bytes(e, 2, shape_proto(v.shape));
    if (off) put_int(e, 4, off);
    put_int(e, 5, size);
    put_fixed32(e, 6, crc_mask(crc32c(v.data, static_cast<size_t>(size))));
    entries.push_back({v.name, e});
  }
  const Bytes index = build_table(entries);

  // ---- GenericModelConfig.json, byte-for-byte what export_generic_config emits (:476-490) ----
  const char* cfg =
      "{\n"
      "    \"inputnames\": [\n"
      "        \"shifu_input_0\"\n"
      "      ],\n"
      "    \"properties\": {\n"
      "         \"algorithm\": \"tensorflow\",\n"
      "         \"tags\": [\"serve\"],\n"
      "         \"outputnames\": \"shifu_output_0\",\n"
      "         \"normtype\": \"ZSCALE\"\n"
      "      }\n"
      "}";

  const std::string dir(export_dir);
  int s = mkdir_p(dir + "/variables");
  if (s != SB_OK) return s;
  if ((s = write_file(dir + "/saved_model.pb", sm)) != SB_OK) return s;
  if ((s = write_file(dir + "/variables/variables.index", index)) != SB_OK) return s;
  if ((s = write_file(dir + "/variables/variables.data-00000-of-00001", data)) != SB_OK) return s;
  return write_file(dir + "/GenericModelConfig.json", cfg);
}

int sb_savedmodel_read(const char* saved_model_dir, const char* input_name, const char* output_name, const char* tag,
                       sb_net_desc* desc_out, int32_t* out_act, float* flat, int64_t flat_cap, int64_t* n_params) {
  if (!saved_model_dir || !input_name || !output_name || !tag || !desc_out || !n_params)
    return set_error(SB_ERR_INVALID, "null argument");
  const std::string dir(saved_model_dir);
  Bytes pb;
  int s = read_file(dir + "/saved_model.pb", &pb);
  if (s != SB_OK) return s;
  std::vector<Field> smf;
  if (!parse_msg(pb.data(), pb.size(), &smf)) return set_error(SB_ERR_FORMAT, "saved_model.pb: not a protobuf");
  std::map<std::string, NodeR> nodes;
  bool found = false;
  for (auto& f : smf) {
    if (f.no != 2 || f.wt != 2) continue;
    std::vector<Field> mg;
    if (!parse_msg(f.p, f.len, &mg)) return set_error(SB_ERR_FORMAT, "saved_model.pb: bad MetaGraphDef");
    bool has_tag = false;
    for (auto& g : mg) {
      if (g.no != 1 || g.wt != 2) continue;
      std::vector<Field> mi;
      if (!parse_msg(g.p, g.len, &mi)) return set_error(SB_ERR_FORMAT, "saved_model.pb: bad MetaInfoDef");
      for (auto& t : mi) if (t.no == 4 && t.wt == 2 && fstr(t) == tag) has_tag = true;
    }
    if (!has_tag) continue;
    for (auto& g : mg) {
      if (g.no != 2 || g.wt != 2) continue;
      std::vector<Field> gd;
      if (!parse_msg(g.p, g.len, &gd)) return set_error(SB_ERR_FORMAT, "saved_model.pb: bad GraphDef");
      for (auto& nf : gd) {
        if (nf.no != 1 || nf.wt != 2) continue;
        std::vector<Field> nd;
        if (!parse_msg(nf.p, nf.len, &nd)) return set_error(SB_ERR_FORMAT, "saved_model.pb: bad NodeDef");
        std::string name;
        NodeR nr;
        for (auto& a : nd) {
          if (a.wt != 2) continue;
          if (a.no == 1) name = fstr(a);
          else if (a.no == 2) nr.op = fstr(a);
          else if (a.no == 3) nr.inputs.push_back(fstr(a));
        }
        nodes[name] = nr;
      }
    }
    found = true;
    break;
  }
  if (!found) return set_error(SB_ERR_FORMAT, "no MetaGraphDef tagged '%s' in %s", tag, dir.c_str());

  // ---- bundle index ----
  Bytes idx;
  if ((s = read_file(dir + "/variables/variables.index", &idx)) != SB_OK) return s;
  std::vector<std::pair<Bytes, Bytes>> ents;
  if ((s = read_table_entries(idx, &ents)) != SB_OK) return s;
  std::map<std::string, BundleEntry> bundle;
  for (auto& kv : ents) {
    if (kv.first.empty()) continue;
    std::vector<Field> ef;
    if (!parse_msg(kv.second.data(), kv.second.size(), &ef)) return set_error(SB_ERR_FORMAT, "bundle: bad entry for %s", kv.first.c_str());
    BundleEntry be;
    for (auto& e : ef) {
      if (e.no == 1 && e.wt == 0) be.dtype = static_cast<int>(e.ival);
      else if (e.no == 2 && e.wt == 2) {
        std::vector<Field> sh;
        if (!parse_msg(e.p, e.len, &sh)) return set_error(SB_ERR_FORMAT, "bundle: bad shape");
        for (auto& dmf : sh) {
          if (dmf.no != 2 || dmf.wt != 2) continue;
          std::vector<Field> dm;
          if (!parse_msg(dmf.p, dmf.len, &dm)) return set_error(SB_ERR_FORMAT, "bundle: bad dim");
          int64_t sz = 0;
          for (auto& x : dm) if (x.no == 1 && x.wt == 0) sz = static_cast<int64_t>(x.ival);
          be.shape.push_back(sz);
        }
      } else if (e.no == 3 && e.wt == 0) { if (e.ival != 0) return set_error(SB_ERR_FORMAT, "bundle: sharded bundles unsupported"); }
      else if (e.no == 4 && e.wt == 0) be.offset = static_cast<int64_t>(e.ival);
      else if (e.no == 5 && e.wt == 0) be.size = static_cast<int64_t>(e.ival);
    }
    bundle[kv.first] = be;
  }

  // ---- graph walk: output -> input ----
  auto node_of = [&](const std::string& n) -> const NodeR* {
    auto it = nodes.find(n);
    return it == nodes.end() ? nullptr : &it->second;
  };
  auto skip_passthrough = [&](std::string n, std::string* out) -> int {
    n = strip_name(n);
    for (int guard = 0; guard < 64; ++guard) {
      const NodeR* nd = node_of(n);
      if (!nd) return set_error(SB_ERR_FORMAT, "graph: node '%s' not found", n.c_str());
      if ((nd->op == "Identity" || nd->op == "StopGradient") && !nd->inputs.empty()) {
        // a variable read (Identity <- VariableV2) is not a pass-through of activations; callers never hit that case
        n = strip_name(nd->inputs[0]);
      } else if (nd->op == "Merge") {
        // Keras dropout K.in_train_phase: Merge(cond/Switch_1 [inference branch], cond/dropout/mul)
        std::string nxt;
        for (auto& in : nd->inputs) {
          const NodeR* c = node_of(strip_name(in));
          if (c && c->op == "Switch" && !c->inputs.empty()) { nxt = strip_name(c->inputs[0]); break; }
        }
        if (nxt.empty()) return set_error(SB_ERR_FORMAT, "graph: unsupported Merge at '%s'", n.c_str());
        n = nxt;
      } else if (nd->op == "Switch" && !nd->inputs.empty()) {
        n = strip_name(nd->inputs[0]);
      } else {
        *out = n;
        return SB_OK;
      }
    }
    return set_error(SB_ERR_FORMAT, "graph: pass-through chain too long at '%s'", n.c_str());
  };
  auto resolve_var = [&](std::string n, std::string* out) -> int {
    n = strip_name(n);
    for (int guard = 0; guard < 8; ++guard) {
      const NodeR* nd = node_of(n);
      if (!nd) return set_error(SB_ERR_FORMAT, "graph: node '%s' not found", n.c_str());
      if (nd->op == "VariableV2" || nd->op == "Variable" || nd->op == "VarHandleOp") { *out = n; return SB_OK; }
      if ((nd->op == "Identity" || nd->op == "ReadVariableOp") && !nd->inputs.empty()) { n = strip_name(nd->inputs[0]); continue; }
      return set_error(SB_ERR_FORMAT, "graph: cannot resolve a variable from '%s' (%s)", n.c_str(), nd->op.c_str());
    }
    return set_error(SB_ERR_FORMAT, "graph: variable chain too long");
  };
  auto is_var_read = [&](const std::string& n) -> bool {
    std::string tmp;
    const NodeR* nd = node_of(strip_name(n));
    if (!nd) return false;
    if (nd->op == "VariableV2" || nd->op == "Variable" || nd->op == "VarHandleOp") return true;
    if ((nd->op == "Identity" || nd->op == "ReadVariableOp") && !nd->inputs.empty()) {
      const NodeR* c = node_of(strip_name(nd->inputs[0]));
      return c && (c->op == "VariableV2" || c->op == "Variable" || c->op == "VarHandleOp");
    }
    return false;
  };

  struct LayerR { std::string w, b; int act; };
  std::vector<LayerR> rl;
  std::string cur;
  if ((s = skip_passthrough(output_name, &cur)) != SB_OK) return s;
  const std::string target = strip_name(input_name);
  if (!node_of(target)) return set_error(SB_ERR_FORMAT, "graph: input '%s' not found", target.c_str());
  for (int guard = 0; cur != target; ++guard) {
    if (guard > 4 * SB_MAX_HIDDEN + 64) return set_error(SB_ERR_FORMAT, "graph: too many layers");
    const NodeR* nd = node_of(cur);
    int act = SB_ACT_NONE;
    if (nd->op == "Sigmoid") act = SB_ACT_SIGMOID;
    else if (nd->op == "Tanh") act = SB_ACT_TANH;
    else if (nd->op == "Relu") act = SB_ACT_RELU;
    else if (nd->op == "LeakyRelu") act = SB_ACT_LEAKYRELU;
    if (act != SB_ACT_NONE) {
      if (nd->inputs.empty()) return set_error(SB_ERR_FORMAT, "graph: activation '%s' has no input", cur.c_str());
      if ((s = skip_passthrough(nd->inputs[0], &cur)) != SB_OK) return s;
      nd = node_of(cur);
    }
    if (!(nd->op == "BiasAdd" || nd->op == "Add" || nd->op == "AddV2") || nd->inputs.size() < 2)
      return set_error(SB_ERR_FORMAT, "graph: expected BiasAdd/Add at '%s', found %s", cur.c_str(), nd->op.c_str());
    std::string mm_in = nd->inputs[0], bias_in = nd->inputs[1];
    if (is_var_read(mm_in)) std::swap(mm_in, bias_in);
    LayerR lr;
    lr.act = act;
    if ((s = resolve_var(bias_in, &lr.b)) != SB_OK) return s;
    std::string mm;
    if ((s = skip_passthrough(mm_in, &mm)) != SB_OK) return s;
    const NodeR* mn = node_of(mm);
    if (mn->op != "MatMul" || mn->inputs.size() < 2)
      return set_error(SB_ERR_FORMAT, "graph: expected MatMul at '%s', found %s", mm.c_str(), mn->op.c_str());
    if ((s = resolve_var(mn->inputs[1], &lr.w)) != SB_OK) return s;
    rl.push_back(lr);
    if ((s = skip_passthrough(mn->inputs[0], &cur)) != SB_OK) return s;
  }
  std::reverse(rl.begin(), rl.end());
  if (rl.size() < 2) return set_error(SB_ERR_FORMAT, "graph: need at least one hidden layer and an output layer");
  if (rl.size() - 1 > SB_MAX_HIDDEN) return set_error(SB_ERR_FORMAT, "graph: %zu hidden layers > SB_MAX_HIDDEN", rl.size() - 1);

  // ---- topology + parameter gather ----
  memset(desc_out, 0, sizeof(*desc_out));
  int64_t total = 0;
  int prev = -1;
  for (size_t l = 0; l < rl.size(); ++l) {
    auto wi = bundle.find(rl[l].w), bi = bundle.find(rl[l].b);
    if (wi == bundle.end() || bi == bundle.end())
      return set_error(SB_ERR_FORMAT, "bundle: variable '%s' / '%s' missing", rl[l].w.c_str(), rl[l].b.c_str());
    const BundleEntry& w = wi->second;
    const BundleEntry& b = bi->second;
    if (w.dtype != DT_FLOAT || b.dtype != DT_FLOAT || w.shape.size() != 2 || b.shape.size() != 1 || b.shape[0] != w.shape[1] ||
        w.size != w.shape[0] * w.shape[1] * 4 || b.size != b.shape[0] * 4)
      return set_error(SB_ERR_FORMAT, "bundle: unexpected dtype/shape for layer %zu", l);
    if (prev >= 0 && w.shape[0] != prev) return set_error(SB_ERR_FORMAT, "graph: layer %zu input width mismatch", l);
    if (l == 0) desc_out->n_features = static_cast<int32_t>(w.shape[0]);
    if (l + 1 < rl.size()) {
      desc_out->hidden[l] = static_cast<int32_t>(w.shape[1]);
      desc_out->acts[l] = rl[l].act;
    } else {
      if (w.shape[1] != 1) return set_error(SB_ERR_FORMAT, "Output now only support single output in inference.");
      if (out_act) *out_act = rl[l].act;
    }
    prev = static_cast<int>(w.shape[1]);
    total += w.shape[0] * w.shape[1] + b.shape[0];
  }
  desc_out->n_hidden = static_cast<int32_t>(rl.size() - 1);
  *n_params = total;
  if (!flat) return SB_OK;
  if (flat_cap < total) return set_error(SB_ERR_INVALID, "flat buffer too small: %lld < %lld", (long long)flat_cap, (long long)total);

  FILE* f = fopen((dir + "/variables/variables.data-00000-of-00001").c_str(), "rb");
  if (!f) return set_error(SB_ERR_IO, "cannot open bundle data file in %s", dir.c_str());
  int64_t off = 0;
  for (size_t l = 0; l < rl.size(); ++l) {
    for (int k = 0; k < 2; ++k) {
      const BundleEntry& e = bundle[k == 0 ? rl[l].w : rl[l].b];
      if (fseek(f, static_cast<long>(e.offset), SEEK_SET) != 0 ||
          fread(flat + off, 1, static_cast<size_t>(e.size), f) != static_cast<size_t>(e.size)) {
        fclose(f);
        return set_error(SB_ERR_IO, "bundle: short read for layer %zu", l);
      }
      off += e.size / 4;
    }
  }
  fclose(f);
  return SB_OK;
}

}  // extern "C"
