// Internal launch interface of the training-side kernels (train.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kge {

struct TrainTables {
  const float* ent0;  // (n_ent, dim)
  const float* ent1;  // second entity plane (ComplEx / RotatE) or nullptr
  const float* rel0;  // (n_rel, dim) or RESCAL (n_rel, dim*dim)
  const float* rel1;  // second relation plane or nullptr
};

struct TrainGrads {
  float* ent0;
  float* ent1;
  float* rel0;
  float* rel1;
};

struct MarginStepParams {
  int model;
  int dim;
  int n_neg;
  float margin;
  long long b;      // positives
  long long n_ent;
  TrainTables tb;
  const int64_t* h;
  const int64_t* t;
  const int64_t* r;
  const int64_t* nh;  // external negatives (b * n_neg, blocks of b) or nullptr -> Philox
  const int64_t* nt;
  const float* probs;  // Bernoulli head-corruption probability per relation (Philox mode)
  uint64_t seed;
  uint64_t offset;
  float* loss;     // 1 float, +=
  float* pos_out;  // optional (b)
  float* neg_out;  // optional (b * n_neg)
  int64_t* nh_out;  // optional
  int64_t* nt_out;
};

cudaError_t launch_score_triples_fwd(int model, int dim, const TrainTables& tb, const int64_t* h,
                                     const int64_t* t, const int64_t* r, int64_t n, float* out,
                                     cudaStream_t st);
cudaError_t launch_score_triples_bwd(int model, int dim, const TrainTables& tb, const TrainGrads& gr,
                                     const int64_t* h, const int64_t* t, const int64_t* r, int64_t n,
                                     const float* gout, cudaStream_t st);
cudaError_t launch_corrupt_batch(const int64_t* h, const int64_t* t, const int64_t* r, int64_t b,
                                 int n_neg, const float* probs, int64_t n_ent, uint64_t seed,
                                 uint64_t offset, int64_t* nh, int64_t* nt, cudaStream_t st);
cudaError_t launch_margin_step_fwd(const MarginStepParams& a, cudaStream_t st);
cudaError_t launch_margin_step_bwd(const MarginStepParams& a, const TrainGrads& gr, const float* gloss,
                                   cudaStream_t st);
cudaError_t launch_margin_loss_fwd(const float* pos, const float* neg, int64_t n, float margin,
                                   float* loss, cudaStream_t st);
cudaError_t launch_pair_loss_fwd(int kind, const float* pos, const float* neg, int64_t n, float* loss,
                                 cudaStream_t st);
cudaError_t launch_pair_loss_bwd(int kind, const float* pos, const float* neg, int64_t n,
                                 const float* gloss, float* gpos, float* gneg, cudaStream_t st);
cudaError_t launch_margin_loss_bwd(const float* pos, const float* neg, int64_t n, float margin,
                                   const float* gloss, float* gpos, float* gneg, cudaStream_t st);

}  // namespace kge
