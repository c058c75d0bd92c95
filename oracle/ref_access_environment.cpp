// TEST INFRASTRUCTURE (oracle/).  src/environment.cpp compiled unmodified + pair-list accessor.
#include "environment.cpp"

extern "C" int ref_pairlist_environment(DerivComputation* c, int* i1, int* i2, int max_edge) {
    auto* p = dynamic_cast<EnvironmentCoverage*>(c);
    if(!p) return -2;
    for(int e=0; e<p->igraph.n_edge && e<max_edge; ++e) { i1[e] = p->igraph.edge_indices1[e]; i2[e] = p->igraph.edge_indices2[e]; }
    return p->igraph.n_edge;
}
