// Host build of pyskani_b200/csrc/gbdt_model.h (tests/test_gbdt_host.py): the very parser and evaluator libskb uses.
#include <cstdio>
#include <cstring>
#include "../../pyskani_b200/csrc/gbdt_model.h"

extern "C" {
// returns 0 and fills out[n_rows]; on a parse error returns 1 and copies the message to err (<= 255 chars)
int gbdt_shim_predict(const char* json, unsigned long long len, const float* rows, unsigned n_rows, unsigned n_features, float* out,
                      unsigned* n_trees, unsigned* n_nodes, char* err) {
    try {
        skb::GbdtHost m = skb::gbdt_parse(json, (size_t)len);
        if (n_trees) *n_trees = (unsigned)m.tree_off.size() - 1;
        if (n_nodes) *n_nodes = (unsigned)m.nodes.size();
        const skb::GbdtView v = m.view();
        for (unsigned i = 0; i < n_rows; i++) {
            float x[skb::GBDT_FEATURES] = {0};
            for (unsigned f = 0; f < n_features && f < skb::GBDT_FEATURES; f++) x[f] = rows[(size_t)i * n_features + f];
            out[i] = skb::gbdt_predict(v, x);
        }
        return 0;
    } catch (const std::exception& e) {
        if (err) { std::strncpy(err, e.what(), 255); err[255] = 0; }
        return 1;
    }
}
}
