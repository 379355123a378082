// ref_shim_cellwise.cpp — the REFERENCE's own cell gather / scatter-add compiled where they lie, and the H.X
// composite assembled from reference-compiled routines only.
//
// TEST INFRASTRUCTURE ONLY (part of oracle/_ref/libdftefe_ref.so, see oracle/Makefile).
//
// basis/FECellWiseDataOperations.h includes basis/Field.h, basis/BasisDataStorage.h and (through them)
// basis/BasisManager.h -> basis/DealiiConversions.h -> <deal.II/...>, which is absent here.  The only thing the
// two functions of the hot path use from those headers is the typedef BasisManager::SizeTypeVector
// (basis/BasisManager.h:59), so the three include guards are pre-defined and that one typedef is supplied; the
// function BODIES compiled below are the reference's (basis/FECellWiseDataOperations.t.cpp:58-153).
#define dftefeField_h
#define dftefeBasisDataStorage_h
#define dftefeBasisManager_h
#include <utils/TypeConfig.h>
#include <utils/MemorySpaceType.h>
#include <utils/MemoryStorage.h>
namespace dftefe
{
  namespace basis
  {
    template <typename ValueType, utils::MemorySpace memorySpace>
    struct BasisManager
    {
      using SizeTypeVector = utils::MemoryStorage<size_type, memorySpace>; // basis/BasisManager.h:59
    };
  } // namespace basis
} // namespace dftefe
#include <basis/FECellWiseDataOperations.h>

#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

using namespace dftefe;
constexpr auto HOST = utils::MemorySpace::HOST;
using CellOps       = basis::FECellWiseDataOperations<double, HOST>;
using SizeVec       = basis::BasisManager<double, HOST>::SizeTypeVector;

// reference-compiled routines exported by ref_shim.cpp (same shared object)
extern "C"
{
  void ref_p2c(double *, unsigned, unsigned, unsigned, const unsigned *, const unsigned *, const unsigned *, const unsigned *,
               unsigned, const double *, const double *);
  void ref_c2p(double *, unsigned, unsigned, unsigned, const unsigned *, const unsigned *, const unsigned *, const unsigned *,
               unsigned, const double *);
  void ref_gemm_strided_var_batched(unsigned, const char *, const char *, const unsigned *, const unsigned *, const unsigned *,
                                    const unsigned *, const unsigned *, const unsigned *, double, const double *, const unsigned *,
                                    const double *, const unsigned *, double, double *, const unsigned *);
  void ref_scale_rows_strided(const double *V, double *cx, unsigned B, unsigned nLocal);
}

static SizeVec
sizes(const unsigned *n, unsigned count)
{
  SizeVec s(count);
  for (unsigned i = 0; i < count; ++i)
    *(s.data() + i) = n[i];
  return s;
}

// the gemmStridedVarBatched call shapes of KohnShamOperatorContextFE.t.cpp:713-760,1155-1175 (H block) and of
// AtomCenterNonLocalOpContextFE.t.cpp:43-153 (projector blocks): m = numVecs, n = nY, k = nX per cell
static void
cell_gemm(unsigned nCells, unsigned B, const unsigned *nX, const unsigned *nY, bool conjTransB, double beta, const double *A,
          const double *Bm, double *Cm)
{
  std::vector<char>     ta(nCells, 'N'), tb(nCells, conjTransB ? 'C' : 'N');
  std::vector<unsigned> m(nCells, B), n(nCells), k(nCells), lda(nCells, B), ldb(nCells), ldc(nCells, B), sa(nCells), sb(nCells),
    sc(nCells);
  for (unsigned i = 0; i < nCells; ++i)
    {
      n[i]   = nY[i];
      k[i]   = nX[i];
      ldb[i] = conjTransB ? n[i] : k[i];
      sa[i]  = B * k[i];
      sb[i]  = k[i] * n[i];
      sc[i]  = B * n[i];
    }
  ref_gemm_strided_var_batched(nCells, ta.data(), tb.data(), sa.data(), sb.data(), sc.data(), m.data(), n.data(), k.data(), 1.0, A,
                               lda.data(), Bm, ldb.data(), beta, Cm, ldc.data());
}

extern "C"
{
  // FECellWiseDataOperations::copyFieldToCellWiseData / addCellWiseDataToFieldData
  void
  ref_gather(const double *x, unsigned B, const unsigned *ids, const unsigned *ncd, unsigned C, double *out)
  {
    CellOps::copyFieldToCellWiseData(x, B, ids, sizes(ncd, C), out);
  }
  void
  ref_scatter_add(const double *in, unsigned B, const unsigned *ids, const unsigned *ncd, unsigned C, double *y)
  {
    CellOps::addCellWiseDataToFieldData(in, B, ids, sizes(ncd, C), y);
  }

  // KohnShamOperatorContextFE::apply on ONE rank (ksdft/KohnShamOperatorContextFE.t.cpp:1313-1443; the ghost
  // exchanges are no-ops without neighbours) with computeAxCellWiseOptimized (:951-1199) and the nonlocal calls of
  // AtomCenterNonLocalOpContextFE (:859-1045).  Only the call sequence and the cell-block loop are restated here:
  // every arithmetic routine invoked is the reference's own, compiled from /root/reference/src.
  void
  ref_hx_apply_serial(double *X, double *Y, unsigned nLocal, unsigned B, unsigned C, const unsigned *ncd, const unsigned *ids,
                      const double *hcell, unsigned nR, const unsigned *rowIds, const unsigned *rowSizes,
                      const unsigned *rowOffsets, const unsigned *colIds, unsigned nnz, const double *colVals,
                      const double *inhom, const unsigned *ncp, const unsigned *pids, const double *cellC, const double *V,
                      unsigned nProjLocal, unsigned cellBlockSize)
  {
    ref_p2c(X, nLocal, B, nR, rowIds, rowSizes, rowOffsets, colIds, nnz, colVals, inhom); // :1352-1353
    std::memset(Y, 0, sizeof(double) * (size_t)nLocal * B);                              // :1372
    const bool          nl = ncp != nullptr;
    size_t              S = 0, SP = 0;
    unsigned            maxDof = 0, maxProj = 0;
    for (unsigned c = 0; c < C; ++c)
      {
        S += ncd[c];
        maxDof = std::max(maxDof, ncd[c]);
        if (nl)
          {
            SP += ncp[c];
            maxProj = std::max(maxProj, ncp[c]);
          }
      }
    std::vector<double> xCell(S * B), yCell((size_t)cellBlockSize * B * maxDof), CX((size_t)nProjLocal * B, 0.0),
      CXCell((size_t)cellBlockSize * B * std::max(maxProj, 1u));
    // first loop: gather + C^H X per cell block (:1031-1075)
    size_t idsOff = 0, pOff = 0, cOff = 0;
    for (unsigned c0 = 0; c0 < C; c0 += cellBlockSize)
      {
        const unsigned nb = std::min(cellBlockSize, C - c0);
        CellOps::copyFieldToCellWiseData(X, B, ids + idsOff, sizes(ncd + c0, nb), xCell.data() + idsOff * B);
        if (nl)
          {
            cell_gemm(nb, B, ncd + c0, ncp + c0, true, 0.0, xCell.data() + idsOff * B, cellC + cOff, CXCell.data());
            CellOps::addCellWiseDataToFieldData(CXCell.data(), B, pids + pOff, sizes(ncp + c0, nb), CX.data());
          }
        for (unsigned i = 0; i < nb; ++i)
          {
            idsOff += ncd[c0 + i];
            if (nl)
              {
                pOff += ncp[c0 + i];
                cOff += (size_t)ncp[c0 + i] * ncd[c0 + i];
              }
          }
      }
    if (nl)
      ref_scale_rows_strided(V, CX.data(), B, nProjLocal); // applyVOnCconjtransX (:963-986)
    // second loop: H block gemm, + C (V C^H X), scatter-add (:1089-1198)
    idsOff = pOff = cOff = 0;
    size_t hOff = 0;
    for (unsigned c0 = 0; c0 < C; c0 += cellBlockSize)
      {
        const unsigned nb = std::min(cellBlockSize, C - c0);
        cell_gemm(nb, B, ncd + c0, ncd + c0, false, 0.0, xCell.data() + idsOff * B, hcell + hOff, yCell.data());
        if (nl)
          {
            CellOps::copyFieldToCellWiseData(CX.data(), B, pids + pOff, sizes(ncp + c0, nb), CXCell.data());
            cell_gemm(nb, B, ncp + c0, ncd + c0, false, 1.0, CXCell.data(), cellC + cOff, yCell.data());
          }
        CellOps::addCellWiseDataToFieldData(yCell.data(), B, ids + idsOff, sizes(ncd + c0, nb), Y);
        for (unsigned i = 0; i < nb; ++i)
          {
            idsOff += ncd[c0 + i];
            hOff += (size_t)ncd[c0 + i] * ncd[c0 + i];
            if (nl)
              {
                pOff += ncp[c0 + i];
                cOff += (size_t)ncp[c0 + i] * ncd[c0 + i];
              }
          }
      }
    ref_c2p(Y, nLocal, B, nR, rowIds, rowSizes, rowOffsets, colIds, nnz, colVals); // :1424-1425
  }

  // The same apply split at its collective point, for a rank of a PARTITIONED run (the caller performs the ghost update
  // before phase A, applyAllReduceOnCconjtransX between the phases and accumulateAddLocallyOwned after phase B, as
  // KohnShamOperatorContextFE::apply does through MPI, :1340-1443).  xCell (S*B doubles) carries the gathered cell data
  // from phase A to phase B like d_XCellWise does in the reference; CX: nProjLocal*B partial sums out of phase A, the
  // reduced and V-scaled coefficients into phase B.
  void
  ref_hx_phase_a(double *X, double *Y, unsigned nLocal, unsigned B, unsigned C, const unsigned *ncd, const unsigned *ids,
                 unsigned nR, const unsigned *rowIds, const unsigned *rowSizes, const unsigned *rowOffsets,
                 const unsigned *colIds, unsigned nnz, const double *colVals, const double *inhom, const unsigned *ncp,
                 const unsigned *pids, const double *cellC, unsigned nProjLocal, unsigned cellBlockSize, double *xCell, double *CX)
  {
    ref_p2c(X, nLocal, B, nR, rowIds, rowSizes, rowOffsets, colIds, nnz, colVals, inhom);
    std::memset(Y, 0, sizeof(double) * (size_t)nLocal * B);
    const bool nl = ncp != nullptr;
    unsigned   maxProj = 0;
    if (nl)
      {
        std::memset(CX, 0, sizeof(double) * (size_t)nProjLocal * B);
        for (unsigned c = 0; c < C; ++c)
          maxProj = std::max(maxProj, ncp[c]);
      }
    std::vector<double> CXCell((size_t)cellBlockSize * B * std::max(maxProj, 1u));
    size_t              idsOff = 0, pOff = 0, cOff = 0;
    for (unsigned c0 = 0; c0 < C; c0 += cellBlockSize)
      {
        const unsigned nb = std::min(cellBlockSize, C - c0);
        CellOps::copyFieldToCellWiseData(X, B, ids + idsOff, sizes(ncd + c0, nb), xCell + idsOff * B);
        if (nl)
          {
            cell_gemm(nb, B, ncd + c0, ncp + c0, true, 0.0, xCell + idsOff * B, cellC + cOff, CXCell.data());
            CellOps::addCellWiseDataToFieldData(CXCell.data(), B, pids + pOff, sizes(ncp + c0, nb), CX);
          }
        for (unsigned i = 0; i < nb; ++i)
          {
            idsOff += ncd[c0 + i];
            if (nl)
              {
                pOff += ncp[c0 + i];
                cOff += (size_t)ncp[c0 + i] * ncd[c0 + i];
              }
          }
      }
  }
  void
  ref_hx_phase_b(double *Y, unsigned nLocal, unsigned B, unsigned C, const unsigned *ncd, const unsigned *ids,
                 const double *hcell, unsigned nR, const unsigned *rowIds, const unsigned *rowSizes,
                 const unsigned *rowOffsets, const unsigned *colIds, unsigned nnz, const double *colVals, const unsigned *ncp,
                 const unsigned *pids, const double *cellC, unsigned cellBlockSize, const double *xCell, const double *CX)
  {
    const bool nl = ncp != nullptr;
    unsigned   maxDof = 0, maxProj = 0;
    for (unsigned c = 0; c < C; ++c)
      {
        maxDof = std::max(maxDof, ncd[c]);
        if (nl)
          maxProj = std::max(maxProj, ncp[c]);
      }
    std::vector<double> yCell((size_t)cellBlockSize * B * maxDof), CXCell((size_t)cellBlockSize * B * std::max(maxProj, 1u));
    size_t              idsOff = 0, pOff = 0, cOff = 0, hOff = 0;
    for (unsigned c0 = 0; c0 < C; c0 += cellBlockSize)
      {
        const unsigned nb = std::min(cellBlockSize, C - c0);
        cell_gemm(nb, B, ncd + c0, ncd + c0, false, 0.0, xCell + idsOff * B, hcell + hOff, yCell.data());
        if (nl)
          {
            CellOps::copyFieldToCellWiseData(CX, B, pids + pOff, sizes(ncp + c0, nb), CXCell.data());
            cell_gemm(nb, B, ncp + c0, ncd + c0, false, 1.0, CXCell.data(), cellC + cOff, yCell.data());
          }
        CellOps::addCellWiseDataToFieldData(yCell.data(), B, ids + idsOff, sizes(ncd + c0, nb), Y);
        for (unsigned i = 0; i < nb; ++i)
          {
            idsOff += ncd[c0 + i];
            hOff += (size_t)ncd[c0 + i] * ncd[c0 + i];
            if (nl)
              {
                pOff += ncp[c0 + i];
                cOff += (size_t)ncp[c0 + i] * ncd[c0 + i];
              }
          }
      }
    ref_c2p(Y, nLocal, B, nR, rowIds, rowSizes, rowOffsets, colIds, nnz, colVals);
  }
}
