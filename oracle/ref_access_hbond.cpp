// TEST INFRASTRUCTURE (oracle/).  src/hbond.cpp compiled unmodified + pair-list accessors (see ref_access_rotamer.cpp).
#include "hbond.cpp"

template <typename G> static int copy_edges(G& g, int* i1, int* i2, int max_edge) {
    for(int e=0; e<g.n_edge && e<max_edge; ++e) { i1[e] = g.edge_indices1[e]; i2[e] = g.edge_indices2[e]; }
    return g.n_edge;
}
extern "C" int ref_pairlist_hbond(DerivComputation* c, int* i1, int* i2, int max_edge) {
    if(auto* p = dynamic_cast<ProteinHBond*>(c))  return copy_edges(p->igraph, i1,i2,max_edge);
    if(auto* p = dynamic_cast<HBondCoverage*>(c)) return copy_edges(p->igraph, i1,i2,max_edge);
    return -2;
}
