#include "prover_impl.cuh"
namespace b2m {
IndexBase* make_index_bls(b2m_srs* srs, int pc, size_t nc, size_t nv, size_t ni, const b2m_matrix* a, const b2m_matrix* b,
                           const b2m_matrix* c) {
  std::unique_ptr<MarlinIndex<FrBls, FqBls>> idx(new MarlinIndex<FrBls, FqBls>(srs, srs->ctx->ntt_bls(), *srs->bls, pc, nc, nv, ni));
  idx->build(a, b, c);
  return idx.release();
}
}  // namespace b2m
