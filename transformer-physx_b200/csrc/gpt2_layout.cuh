// Packed-weight layout of the decoder stack (floats).  Shared by the packer and every kernel.
// Reference parameter shapes: trphysx/transformer/phys_transformer_gpt2.py:126-138,
// attention.py:69-70 (c_attn [E,3E], c_proj [E,E]), :39-41 (mlp c_fc [E,4E], c_proj [4E,E]).
#pragma once

namespace trpx {

struct Gpt2Layout {
    int E, n_layer, n_ctx, n_head, hd;
    // per-layer offsets (relative to the layer base)
    int ln1w, ln1b, wqkv, bqkv, wproj, bproj, ln2w, ln2b, wfc, bfc, wpr, bpr, layer_stride;
    // global offsets (relative to the blob base)
    int lnfw, lnfb, wf, bf, total;

    __host__ __device__ static Gpt2Layout make(int E, int n_layer, int n_ctx, int n_head) {
        Gpt2Layout L;
        L.E = E; L.n_layer = n_layer; L.n_ctx = n_ctx; L.n_head = n_head; L.hd = E / n_head;
        int o = 0;
        L.ln1w = o; o += E;
        L.ln1b = o; o += E;
        L.wqkv = o; o += E * 3 * E;
        L.bqkv = o; o += 3 * E;
        L.wproj = o; o += E * E;
        L.bproj = o; o += E;
        L.ln2w = o; o += E;
        L.ln2b = o; o += E;
        L.wfc = o; o += E * 4 * E;
        L.bfc = o; o += 4 * E;
        L.wpr = o; o += 4 * E * E;
        L.bpr = o; o += E;
        L.layer_stride = o;                  // 12 E^2 + 13 E
        int g = n_layer * L.layer_stride;
        L.lnfw = g; g += E;
        L.lnfb = g; g += E;
        L.wf = g; g += E * E;                // stored TRANSPOSED: [in][out]
        L.bf = g; g += E;
        L.total = g;
        return L;
    }
};

}  // namespace trpx
