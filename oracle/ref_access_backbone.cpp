// TEST INFRASTRUCTURE (oracle/).  src/backbone_steric.cpp compiled unmodified + residue-pair-list accessor.
#include "backbone_steric.cpp"

extern "C" int ref_pairlist_backbone(DerivComputation* c, int* i1, int* i2, int max_edge) {
    auto* p = dynamic_cast<BackbonePairs*>(c);
    if(!p) return -2;
    for(int e=0; e<p->pairlist.n_edge && e<max_edge; ++e) { i1[e] = p->pairlist.edge_indices1[e]; i2[e] = p->pairlist.edge_indices2[e]; }
    return p->pairlist.n_edge;
}
