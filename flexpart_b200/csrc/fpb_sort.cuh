// fpb_sort.cuh -- re-ordering of the device-resident particle rows by
// meteorological grid cell (level-major: key = met array index of the cell
// under the particle, below the turbulence regime of the coming step), for
// gather locality and warp convergence.
// slot[row] / row_of_slot[slot] keep the caller's slot indices stable.
#pragma once
#include "fpb_device.cuh"
#include "fpb_scatter.cuh"

// build keys (dead rows sort last), returns the number of live rows via *d_nlive.
// met != null: met[0..1] = the fields bracketing c.itime; the turbulence regime of
// the coming step is put above the cell bits of the key.  The cell part is
// (level, jy, ix) level-major, coarsened to tiles of 2^k cells when the exact index
// would need a 4th 8-bit radix pass; sortk_key_bits() = number of key bits to sort.
void sortk_build_keys(const DevCfg &c, const DevParticles &p, const float *height, int nrows,
                      unsigned *keys, unsigned *ids, unsigned *d_nlive, cudaStream_t st,
                      const DevMetSlot *met);
int sortk_key_layout_bits(const DevCfg &c);
int sortk_key_bits(const DevCfg &c, bool regime);
// dst row i := src row ids[i] for every particle array (incl. slot);
// row_of_slot[slot] := base + i
void sortk_permute(const DevParticles &src, const DevParticles &dst, const unsigned *ids,
                   int nrows, int nspec, cudaStream_t st, int32_t *row_of_slot, int base = 0);
// the same through packed records (`records`: sortk_packed_bytes() bytes of scratch; 0 = not available for
// this nspec): two coalesced passes instead of one gather that fetches a sector per value
size_t sortk_packed_bytes(int nrows, int nspec);
void sortk_permute_packed(const DevParticles &src, const DevParticles &dst, const unsigned *ids, int nrows, int nspec,
                          cudaStream_t st, int32_t *row_of_slot, void *records, int base = 0);
void sortk_invert(const int32_t *slot, int32_t *row_of_slot, int nrows, cudaStream_t st, int base = 0);
// staging row slot[i] <- row i of the view `rows` (arrays the particle loop writes only)
void sortk_scatter_back(const DevParticles &rows, const DevParticles &stg, int count, int nspec,
                        cudaStream_t st, bool scav = false);
void sortk_iota(int32_t *a, int n, cudaStream_t st);
// staging (slot order, rows [first,first+count)) <-> device rows
void sortk_gather_to_staging(const DevParticles &rows, const DevParticles &stg,
                             const int32_t *row_of_slot, int first, int count, int nspec,
                             cudaStream_t st);
void sortk_scatter_from_staging(const DevParticles &stg, const DevParticles &rows,
                                const int32_t *row_of_slot, int first, int count, int nspec,
                                bool have_split, bool have_scav, cudaStream_t st);
