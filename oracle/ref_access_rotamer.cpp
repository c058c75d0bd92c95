// TEST INFRASTRUCTURE (oracle/).  Compiles the reference's src/rotamer.cpp UNMODIFIED (textual include from
// where it lies) and adds read-only accessors for state that the reference keeps in file-local types:
// the rotamer bead-pair list, BP node marginals and the BP iteration count.
#include "rotamer.cpp"

extern "C" {
// returns n_edge (or -2 if `c` is not a rotamer node); fills at most max_edge entries, reference emission order
int ref_pairlist_rotamer(DerivComputation* c, int* i1, int* i2, int max_edge) {
    auto* r = dynamic_cast<RotamerSidechain<preferred_bead_type>*>(c);
    if(!r) return -2;
    int n = r->igraph.n_edge;
    for(int e=0; e<n && e<max_edge; ++e) { i1[e] = r->igraph.edge_indices1[e]; i2[e] = r->igraph.edge_indices2[e]; }
    return n;
}

// per-bead node marginal (cur_belief after solve), in bead order; returns n_bead or -2
int ref_rotamer_bead_marginals(DerivComputation* c, float* out, int max_bead) {
    auto* r = dynamic_cast<RotamerSidechain<preferred_bead_type>*>(c);
    if(!r) return -2;
    int n = r->igraph.n_elem1;
    for(int b=0; b<n && b<max_bead; ++b) {
        unsigned id = r->igraph.id1[b];
        unsigned sel = (1u<<n_bit_rotamer)-1u;
        unsigned rot = id & sel; id >>= n_bit_rotamer;
        unsigned n_rot = id & sel; id >>= n_bit_rotamer;
        out[b] = r->node_holders_matrix[n_rot]->cur_belief(rot,id);
    }
    return n;
}

// re-runs fill+solve on the current inputs and reports (iterations, final max deviation, n residue-pair edges
// 33/36/66/11/13/16); the reference discards the iteration count inside compute_value (rotamer.cpp:782-785)
int ref_rotamer_solve_stats(DerivComputation* c, float* out8) {
    auto* r = dynamic_cast<RotamerSidechain<preferred_bead_type>*>(c);
    if(!r) return -2;
    r->fill_holders();
    auto res = r->solve_for_marginals();
    out8[0] = res.first; out8[1] = res.second;
    out8[2] = r->edges33.nodes_to_edge.n_edge; out8[3] = r->edges36.nodes_to_edge.n_edge;
    out8[4] = r->edges66.nodes_to_edge.n_edge; out8[5] = r->edges11.nodes_to_edge.n_edge;
    out8[6] = r->edges13.nodes_to_edge.n_edge; out8[7] = r->edges16.nodes_to_edge.n_edge;
    return 0;
}
}
