/* b200_synth.h — synthetic graph generators of the HARNESS (libb200synth.so; host code only).
 *
 * Not part of the product library: tests, the bench (both arms) and the example driver use them;
 * the reference arm of bench.py therefore loads only this library and oracle/_ref/libref.so.
 * The reference has no generators (its drivers read SNAP / MatrixMarket files that are not in
 * the repository); BASELINE.json names the three families, SURVEY.md §8(d) fixes their
 * parameters.  Blocks are plain malloc(): release them with free() (or b200_host_free). */
#ifndef B200_SYNTH_H_
#define B200_SYNTH_H_
#ifdef __cplusplus
extern "C" {
#endif

/* * All return a malloc()'d int CSR with rmclInit semantics (nlibs/qrmcl.cc:126-134): self loop
 * on every vertex, sorted unique columns, values 1/rowcount.  Release with free(). */
int b200_synth_rmat(int scale, int edge_factor, unsigned long long seed, int symmetrise,
                    int* rows, int** IA, int** JA, double** A, long long* nnz);
int b200_synth_stencil27(int gx, int gy, int gz, int* rows, int** IA, int** JA, double** A,
                         long long* nnz);
int b200_synth_planted(int n, int nblocks, int intra, int inter, unsigned long long seed,
                       int* rows, int** IA, int** JA, double** A, long long* nnz,
                       int** labels);
#ifdef __cplusplus
}
#endif
#endif /* B200_SYNTH_H_ */
