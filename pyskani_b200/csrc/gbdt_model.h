// gbdt_model.h — loader and evaluator of a gbdt-rs regression ensemble (the "learned ANI" model of skani).
//
// Replaces skani::regression::get_model + GBDT::predict as used through the reference at lib.rs:611-614 and inside
// skani::chain::chain_seeds (lib.rs:652-653).  skani embeds its models as serde_json dumps of the crate `gbdt` 0.1.3
// (Cargo.lock:541-542); those weights are not part of the reference's sources, so the model comes from a file the user
// supplies in the same format:
//
//   GBDT         { "conf": Config, "trees": [DecisionTree], "bias": f32 }
//   Config       { "feature_size", "max_depth", "iterations", "shrinkage", "loss", "initial_guess_enabled", ... }
//   DecisionTree { "tree": { "tree": [BinaryTreeNode] }, ... }
//   BinaryTreeNode { "value": DTNode, "index", "left", "right" }      (left/right == 0: no child; node 0 is the root)
//   DTNode       { "feature_index", "feature_value", "pred", "missing" (-1 left, 0 stop, 1 right), "is_leaf" }
//
// Prediction as gbdt-rs computes it, in f32 (its ValueType): p = bias; for each of the first `iterations` trees
// p += shrinkage * tree(x), where a tree walks from the root: leaf -> pred; feature == f32::MIN (unknown) -> by `missing`;
// feature < feature_value -> left, else right.  Multiplication and addition are separate f32 roundings (Rust does not
// fuse them), and the trees are added in order: gbdt_predict below does exactly that on the host and on the device.
//
// No CUDA types here: the parser is also compiled for the host by tests/host_shim.
#pragma once
#include <stdint.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#if defined(__CUDACC__)
#define GBDT_HD __host__ __device__ __forceinline__
#else
#define GBDT_HD inline
#endif

namespace skb {

constexpr uint32_t GBDT_FEATURES = 10;      // length of the feature vector built per pair (see FEATURES in DESIGN.md)
constexpr float GBDT_UNKNOWN = -3.40282347e+38f;   // gbdt-rs VALUE_TYPE_UNKNOWN = f32::MIN

struct GbdtNode {            // 20 bytes
    float threshold;         // DTNode.feature_value
    float pred;              // DTNode.pred
    uint32_t feature;        // DTNode.feature_index
    uint32_t left, right;    // indices inside the tree's node range; 0 = none
    int32_t flags;           // bit 0: is_leaf; bits 8..15: missing + 1 (0 left, 1 stop, 2 right)
};

struct GbdtView {            // what the evaluator reads (host or device pointers)
    const GbdtNode* nodes;
    const uint32_t* tree_off;   // [n_trees + 1] first node of every tree
    uint32_t n_trees;           // = min(conf.iterations, trees.len())
    uint32_t n_features;        // conf.feature_size
    float bias, shrinkage;
};

GBDT_HD float gbdt_tree(const GbdtView& m, uint32_t t, const float* x) {
    const GbdtNode* nd = m.nodes + m.tree_off[t];
    const uint32_t n = m.tree_off[t + 1] - m.tree_off[t];
    uint32_t i = 0;
    for (uint32_t guard = 0; guard <= n; guard++) {        // a well-formed tree ends long before; the guard stops cycles
        const GbdtNode nn = nd[i];
        if (nn.flags & 1) return nn.pred;
        const float v = x[nn.feature];
        uint32_t next;
        if (v == GBDT_UNKNOWN) {
            const int miss = ((nn.flags >> 8) & 0xff) - 1;
            if (miss == 0) return nn.pred;
            next = miss < 0 ? nn.left : nn.right;
        } else {
            next = v < nn.threshold ? nn.left : nn.right;
        }
        if (next == 0 || next >= n) return nn.pred;         // gbdt-rs would panic on a missing child; stop at the node instead
        i = next;
    }
    return nd[i].pred;
}

#if defined(__CUDA_ARCH__)
#define GBDT_MUL(a, b) __fmul_rn((a), (b))
#define GBDT_ADD(a, b) __fadd_rn((a), (b))
#else
#define GBDT_MUL(a, b) ((a) * (b))
#define GBDT_ADD(a, b) ((a) + (b))
#endif

// whole ensemble, trees added in order (one thread); the host build must not contract a*b+c (compile with -ffp-contract=off
// or rely on the volatile below)
GBDT_HD float gbdt_predict(const GbdtView& m, const float* x) {
    float p = m.bias;
    for (uint32_t t = 0; t < m.n_trees; t++) {
#if defined(__CUDA_ARCH__)
        p = GBDT_ADD(p, GBDT_MUL(m.shrinkage, gbdt_tree(m, t, x)));
#else
        volatile float prod = m.shrinkage * gbdt_tree(m, t, x);
        p = p + prod;
#endif
    }
    return p;
}

// ------------------------------------------------------------------------------------------------ JSON (host)
namespace gbdt_json {

struct Value;
using ValuePtr = std::shared_ptr<Value>;
struct Value {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;                 // Str, and the literal text of a Num (for exact f32 parsing)
    std::vector<ValuePtr> arr;
    std::map<std::string, ValuePtr> obj;
    const Value& at(const char* key) const {
        auto it = obj.find(key);
        if (kind != Obj || it == obj.end()) throw std::runtime_error(std::string("gbdt model: missing field \"") + key + "\"");
        return *it->second;
    }
    bool has(const char* key) const { return kind == Obj && obj.count(key); }
};

struct Parser {
    const char* p; const char* end; int depth = 0;
    [[noreturn]] void fail(const char* what) const { throw std::runtime_error(std::string("gbdt model: invalid JSON (") + what + ")"); }
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
    ValuePtr parse() {
        if (++depth > 64) fail("nesting too deep");
        ws();
        if (p >= end) fail("unexpected end");
        auto v = std::make_shared<Value>();
        const char c = *p;
        if (c == '{') {
            v->kind = Value::Obj; p++; ws();
            if (p < end && *p == '}') { p++; depth--; return v; }
            while (true) {
                ws();
                if (p >= end || *p != '"') fail("expected a key");
                std::string key = string();
                ws();
                if (p >= end || *p != ':') fail("expected ':'");
                p++;
                v->obj[key] = parse();
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == '}') { p++; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v->kind = Value::Arr; p++; ws();
            if (p < end && *p == ']') { p++; depth--; return v; }
            while (true) {
                v->arr.push_back(parse());
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == ']') { p++; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v->kind = Value::Str; v->str = string();
        } else if (c == 't' && end - p >= 4 && !std::strncmp(p, "true", 4)) { v->kind = Value::Bool; v->b = true; p += 4; }
        else if (c == 'f' && end - p >= 5 && !std::strncmp(p, "false", 5)) { v->kind = Value::Bool; v->b = false; p += 5; }
        else if (c == 'n' && end - p >= 4 && !std::strncmp(p, "null", 4)) { v->kind = Value::Null; p += 4; }
        else if (c == '-' || (c >= '0' && c <= '9')) {
            const char* s = p;
            if (*p == '-') p++;
            while (p < end && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '+' || *p == '-')) p++;
            v->kind = Value::Num; v->str.assign(s, p - s);
            char* e2 = nullptr;
            v->num = std::strtod(v->str.c_str(), &e2);
            if (e2 == v->str.c_str() || *e2) fail("bad number");
        } else fail("unexpected character");
        depth--;
        return v;
    }
    std::string string() {
        std::string out;
        p++;                                   // opening quote
        while (p < end && *p != '"') {
            if (*p == '\\') {
                if (++p >= end) fail("bad escape");
                switch (*p) {
                    case 'n': out.push_back('\n'); break; case 't': out.push_back('\t'); break;
                    case 'r': out.push_back('\r'); break; case 'b': out.push_back('\b'); break;
                    case 'f': out.push_back('\f'); break;
                    case 'u': if (end - p < 5) fail("bad \\u escape"); out.push_back('?'); p += 4; break;
                    default: out.push_back(*p);
                }
                p++;
            } else out.push_back(*p++);
        }
        if (p >= end) fail("unterminated string");
        p++;
        return out;
    }
};

// serde_json prints an f32 with the shortest digits that round-trip as f32; strtof gives that f32 back exactly
inline float as_f32(const Value& v) {
    if (v.kind != Value::Num) throw std::runtime_error("gbdt model: expected a number");
    return std::strtof(v.str.c_str(), nullptr);
}
inline int64_t as_int(const Value& v) {
    if (v.kind != Value::Num) throw std::runtime_error("gbdt model: expected an integer");
    return (int64_t)std::llround(v.num);
}

}  // namespace gbdt_json

struct GbdtHost {
    std::vector<GbdtNode> nodes;
    std::vector<uint32_t> tree_off;
    uint32_t n_features = 0;
    float bias = 0.f, shrinkage = 1.f;
    std::string loss;
    GbdtView view() const { return GbdtView{nodes.data(), tree_off.data(), (uint32_t)tree_off.size() - 1, n_features, bias, shrinkage}; }
};

// Parses the serde_json dump of a gbdt-rs `GBDT`.  Throws std::runtime_error with a message naming what is wrong.
inline GbdtHost gbdt_parse(const char* text, size_t len) {
    using namespace gbdt_json;
    Parser ps{text, text + len};
    ValuePtr root = ps.parse();
    ps.ws();
    if (ps.p != ps.end) ps.fail("trailing characters");
    const Value& conf = root->at("conf");
    GbdtHost m;
    m.n_features = (uint32_t)as_int(conf.at("feature_size"));
    m.shrinkage = as_f32(conf.at("shrinkage"));
    m.bias = root->has("bias") ? as_f32(root->at("bias")) : 0.f;
    if (conf.has("loss")) {
        const Value& l = conf.at("loss");
        m.loss = l.kind == Value::Str ? l.str : "";
    }
    // only regression losses predict the raw sum (GBDT::predict applies a sigmoid for the logistic ones)
    if (!m.loss.empty() && m.loss != "SquaredError" && m.loss != "LAD" && m.loss != "RegLinear")
        throw std::runtime_error("gbdt model: loss \"" + m.loss + "\" is not a regression loss");
    if (conf.has("initial_guess_enabled") && conf.at("initial_guess_enabled").kind == Value::Bool && conf.at("initial_guess_enabled").b)
        throw std::runtime_error("gbdt model: initial_guess_enabled models are not supported");
    if (m.n_features == 0 || m.n_features > GBDT_FEATURES)
        throw std::runtime_error("gbdt model: feature_size must be in 1.." + std::to_string(GBDT_FEATURES));
    const Value& trees = root->at("trees");
    if (trees.kind != Value::Arr) throw std::runtime_error("gbdt model: \"trees\" is not an array");
    size_t use = trees.arr.size();
    if (conf.has("iterations")) use = std::min<size_t>(use, (size_t)std::max<int64_t>(0, as_int(conf.at("iterations"))));
    m.tree_off.push_back(0);
    for (size_t t = 0; t < use; t++) {
        const Value& nodes = trees.arr[t]->at("tree").at("tree");
        if (nodes.kind != Value::Arr || nodes.arr.empty()) throw std::runtime_error("gbdt model: empty tree");
        const size_t n = nodes.arr.size();
        for (size_t i = 0; i < n; i++) {
            const Value& bn = *nodes.arr[i];
            const Value& dn = bn.at("value");
            GbdtNode g{};
            g.threshold = as_f32(dn.at("feature_value"));
            g.pred = as_f32(dn.at("pred"));
            const int64_t fi = as_int(dn.at("feature_index"));
            const bool leaf = dn.at("is_leaf").kind == Value::Bool && dn.at("is_leaf").b;
            const int64_t miss = dn.has("missing") ? as_int(dn.at("missing")) : 0;
            const int64_t l = as_int(bn.at("left")), r = as_int(bn.at("right"));
            if (bn.has("index") && as_int(bn.at("index")) != (int64_t)i) throw std::runtime_error("gbdt model: node index does not match its position");
            if (l < 0 || r < 0 || (size_t)l >= n || (size_t)r >= n) throw std::runtime_error("gbdt model: child index out of range");
            if (!leaf && (fi < 0 || fi >= (int64_t)m.n_features)) throw std::runtime_error("gbdt model: feature_index out of range");
            if (miss < -1 || miss > 1) throw std::runtime_error("gbdt model: \"missing\" must be -1, 0 or 1");
            g.feature = leaf ? 0u : (uint32_t)fi; g.left = (uint32_t)l; g.right = (uint32_t)r;
            g.flags = (leaf ? 1 : 0) | (int32_t)((miss + 1) << 8);
            m.nodes.push_back(g);
        }
        m.tree_off.push_back((uint32_t)m.nodes.size());
    }
    return m;
}

}  // namespace skb
