// HotPath.h — C++17 host-side mirror of the reference's operator-context API for the H.X / Chebyshev /
// Rayleigh-Ritz hot path, implemented on top of the C ABI of libhxb200 (include/hxb200.h).
//
// Names, argument meaning and error behaviour follow the reference so that its call sites compile against
// these classes unchanged for MemorySpace::DEVICE:
//   dftefe::linearAlgebra::OperatorContext<double,double,DEVICE>::apply(X, Y, updateGhostX, updateGhostY)
//                                                   (reference src/linearAlgebra/OperatorContext.h:48-111)
//   dftefe::linearAlgebra::MultiVector<double,DEVICE> (src/linearAlgebra/MultiVector.h:134-508)
//   dftefe::ksdft::KohnShamOperatorContextFE          (src/ksdft/KohnShamOperatorContextFE.h:57-157)
//   dftefe::basis::{CFEOverlapInverseOpContextGLL, OEFEAtomBlockOverlapInvOpContextGLL,
//                   OrthoEFEOverlapOperatorContext}   (src/basis/*OpContext*.h, apply only)
//   dftefe::linearAlgebra::{ChebyshevFilter, ResidualChebyshevFilterGEP}
//                                                   (src/linearAlgebra/ChebyshevFilter.h:54-113)
//   computeXTransOpX / subspaceRotation              (src/linearAlgebra/RayleighRitzEigenSolver.t.cpp:685-844,
//                                                    src/linearAlgebra/ElpaScalapackOperations.t.cpp:185-335)
// The reference's FEBasisManager / ConstraintsLocal / MPIPatternP2P need deal.II and MPI; what the hot path
// consumes of them is their flat index arrays, carried here by basis::FEBasisManagerArrays (INTEGRATION.md
// lists the getter each field is filled from).  Errors surface as utils::HxException, thrown the way the
// reference's utils::throwException does (src/utils/Exceptions.h:128-132).  There is no CPU fallback.
#ifndef DFTEFE_B200_HOTPATH_H
#define DFTEFE_B200_HOTPATH_H

#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../../include/hxb200.h"

namespace dftefe
{
  using size_type        = unsigned int;  // src/utils/TypeConfig.h:8
  using global_size_type = unsigned long; // src/utils/TypeConfig.h:9

  namespace utils
  {
    enum class MemorySpace
    {
      HOST,
      HOST_PINNED,
      DEVICE
    }; // src/utils/MemorySpaceType.h:36-41

    class HxException : public std::runtime_error
    {
    public:
      HxException(int code, const std::string &msg)
        : std::runtime_error(msg)
        , d_code(code)
      {}
      int
      code() const
      {
        return d_code;
      }

    private:
      int d_code;
    };

    inline void
    throwException(bool condition, const std::string &msg = "")
    {
      if (!condition)
        throw HxException(HX_ERR_INVALID, msg);
    }

    inline void
    hxCheck(int rc)
    {
      if (rc != HX_OK)
        throw HxException(rc, std::string("libhxb200: ") + hx_last_error());
    }
  } // namespace utils

  namespace utils
  {
    namespace mpi
    {
      // the getters of utils::mpi::MPIPatternP2P the ghost communicator consumes (src/utils/MPIPatternP2P.h:425-459)
      struct MPIPatternP2PArrays
      {
        size_type              localOwnedSize = 0, localGhostSize = 0;
        std::vector<size_type> ghostProcIds, ghostLocalIndicesRanges, ghostLocalIndicesForGhostProcs;
        std::vector<size_type> targetProcIds, numOwnedIndicesForTargetProcs, ownedLocalIndicesForTargetProcs;

        hx_halo_desc
        desc() const
        {
          hx_halo_desc h;
          h.n_owned                     = localOwnedSize;
          h.n_ghost                     = localGhostSize;
          h.n_ghost_procs               = (uint32_t)ghostProcIds.size();
          h.ghost_proc_ids              = ghostProcIds.data();
          h.ghost_ranges                = ghostLocalIndicesRanges.data();
          h.ghost_local_ids             = ghostLocalIndicesForGhostProcs.data();
          h.n_target_procs              = (uint32_t)targetProcIds.size();
          h.target_proc_ids             = targetProcIds.data();
          h.num_owned_for_target        = numOwnedIndicesForTargetProcs.data();
          h.owned_local_ids_for_targets = ownedLocalIndicesForTargetProcs.data();
          return h;
        }
      };
    } // namespace mpi
  }   // namespace utils

  namespace basis
  {
    // What FEBasisManager + ConstraintsLocal hand to the hot path, as flat arrays
    // (src/basis/FEBasisManager.h:143-160, FEBasisManager.t.cpp:414-428,545-557;
    //  src/basis/CFEConstraintsLocalDealii.t.cpp:286-462).
    struct FEBasisManagerArrays
    {
      int                             rank = 0, nRanks = 1;
      utils::mpi::MPIPatternP2PArrays mpiPatternP2P;
      size_type                       nLocallyOwnedClassicalDofs = 0;
      std::vector<size_type>          nLocallyOwnedCellDofs;       // per cell
      std::vector<size_type>          locallyOwnedCellLocalDofIds; // concatenated
      std::vector<size_type>          rowConstraintsIdsLocal, rowConstraintsSizes, columnConstraintsAccumulated,
        columnConstraintsIdsLocal;
      std::vector<double> columnConstraintsValues, constraintsInhomogenities;
    };

    // AtomCenterNonLocalOpContextFE's apply-time data (src/basis/AtomCenterNonLocalOpContextFE.t.cpp:478-494,595-619)
    struct AtomCenterNonLocalArrays
    {
      utils::mpi::MPIPatternP2PArrays mpiPatternP2PProj;
      std::vector<size_type>          numProjsInCells, locallyOwnedCellLocalProjectorIds;
      std::vector<double>             cellWiseC, V;
    };
  } // namespace basis

  namespace linearAlgebra
  {
    // One per GPU / MPI rank: owns the device copies of the index maps, the ghost communicator buffers and the
    // stream (the role MPIPatternP2P + MPICommunicatorP2P + LinAlgOpContext<DEVICE> play together).
    class DeviceContext
    {
    public:
      DeviceContext(const basis::FEBasisManagerArrays &fe, size_type maxWaveFnBatch, void *cudaStream = nullptr)
        : d_nLocal(fe.mpiPatternP2P.localOwnedSize + fe.mpiPatternP2P.localGhostSize)
        , d_nOwned(fe.mpiPatternP2P.localOwnedSize)
        , d_maxBlock(maxWaveFnBatch)
      {
        hx_mesh_desc m;
        std::memset(&m, 0, sizeof(m));
        m.struct_size       = sizeof(m);
        m.rank              = fe.rank;
        m.nranks            = fe.nRanks;
        m.halo              = fe.mpiPatternP2P.desc();
        m.n_owned_classical = fe.nLocallyOwnedClassicalDofs;
        m.n_cells           = (uint32_t)fe.nLocallyOwnedCellDofs.size();
        m.num_cell_dofs     = fe.nLocallyOwnedCellDofs.data();
        m.cell_local_ids    = fe.locallyOwnedCellLocalDofIds.data();
        m.n_constraint_rows = (uint32_t)fe.rowConstraintsIdsLocal.size();
        m.row_ids           = fe.rowConstraintsIdsLocal.data();
        m.row_sizes         = fe.rowConstraintsSizes.data();
        m.row_offsets       = fe.columnConstraintsAccumulated.data();
        m.col_ids           = fe.columnConstraintsIdsLocal.data();
        m.col_vals          = fe.columnConstraintsValues.data();
        m.inhom             = fe.constraintsInhomogenities.data();
        m.max_block         = maxWaveFnBatch;
        d_S2                = 0;
        for (size_type n : fe.nLocallyOwnedCellDofs)
          d_S2 += (size_t)n * n;
        utils::hxCheck(hx_plan_create(&d_plan, &m, cudaStream));
      }
      ~DeviceContext() { hx_plan_destroy(d_plan); }
      DeviceContext(const DeviceContext &) = delete;
      DeviceContext &
      operator=(const DeviceContext &) = delete;

      // multi-GPU: rank 0 calls uniqueId(), the application broadcasts the 128 bytes (MPI_Bcast), all attach
      static std::vector<char>
      uniqueId()
      {
        std::vector<char> id(128);
        utils::hxCheck(hx_comm_unique_id(id.data()));
        return id;
      }
      void
      attachCommunicator(const std::vector<char> &id)
      {
        utils::throwException(id.size() == 128, "communicator id must be 128 bytes");
        utils::hxCheck(hx_plan_attach_comm(d_plan, id.data()));
      }
      void
      synchronize() const
      {
        utils::hxCheck(hx_plan_synchronize(d_plan));
      }
      hx_plan *
      plan() const
      {
        return d_plan;
      }
      size_type
      localSize() const
      {
        return d_nLocal;
      }
      size_type
      locallyOwnedSize() const
      {
        return d_nOwned;
      }
      size_type
      maxBlock() const
      {
        return d_maxBlock;
      }
      size_t
      cellMatrixSize() const
      {
        return d_S2;
      }

    private:
      hx_plan * d_plan = nullptr;
      size_type d_nLocal, d_nOwned, d_maxBlock;
      size_t    d_S2;
    };

    template <typename ValueType, utils::MemorySpace memorySpace>
    class MultiVector;

    // MultiVector<double, DEVICE>: data()[iDof*numVectors + iVec], owned rows then ghost rows
    // (src/linearAlgebra/MultiVector.h:134-160).
    template <>
    class MultiVector<double, utils::MemorySpace::DEVICE>
    {
    public:
      using value_type = double;
      MultiVector(std::shared_ptr<const DeviceContext> ctx, size_type numVectors, double initVal = 0.0)
        : d_ctx(std::move(ctx))
        , d_numVectors(numVectors)
      {
        utils::throwException(numVectors >= 1 && numVectors <= d_ctx->maxBlock(),
                              "MultiVector: numVectors outside [1, maxWaveFnBatch]");
        utils::hxCheck(hx_device_alloc((void **)&d_data, bytes()));
        setValue(initVal);
      }
      MultiVector(const MultiVector &u)
        : MultiVector(u.d_ctx, u.d_numVectors)
      {
        std::vector<double> tmp(u.localSize() * (size_t)d_numVectors);
        u.copyTo(tmp.data());
        copyFrom(tmp.data());
      }
      MultiVector(MultiVector &&u) noexcept
        : d_ctx(std::move(u.d_ctx))
        , d_numVectors(u.d_numVectors)
        , d_data(u.d_data)
      {
        u.d_data = nullptr;
      }
      ~MultiVector()
      {
        if (d_data)
          hx_device_free(d_data);
      }
      double *
      data()
      {
        return d_data;
      }
      const double *
      data() const
      {
        return d_data;
      }
      double *
      begin()
      {
        return d_data;
      }
      double *
      end()
      {
        return d_data + (size_t)localSize() * d_numVectors;
      }
      void
      setValue(const double val)
      {
        if (val == 0.0)
          utils::hxCheck(hx_memset_zero(d_data, bytes()));
        else
          {
            std::vector<double> tmp((size_t)localSize() * d_numVectors, val);
            copyFrom(tmp.data());
          }
      }
      // host <-> device transfer of the whole local block (utils::MemoryTransfer<DEVICE,HOST> in the reference)
      void
      copyFrom(const double *host)
      {
        utils::hxCheck(hx_memcpy_h2d(d_data, host, bytes()));
      }
      void
      copyTo(double *host) const
      {
        utils::hxCheck(hx_memcpy_d2h(host, d_data, bytes()));
      }
      std::vector<double>
      l2Norms() const
      {
        std::vector<double> out(d_numVectors);
        utils::hxCheck(hx_l2_norms(d_ctx->plan(), d_data, d_numVectors, out.data()));
        return out;
      }
      size_type
      getNumberComponents() const
      {
        return d_numVectors;
      }
      size_type
      numVectors() const
      {
        return d_numVectors;
      }
      void
      updateGhostValues(const size_type /*communicationChannel*/ = 0)
      {
        utils::hxCheck(hx_update_ghost_values(d_ctx->plan(), d_data, d_numVectors));
      }
      void
      accumulateAddLocallyOwned(const size_type /*communicationChannel*/ = 0)
      {
        utils::hxCheck(hx_accumulate_add_locally_owned(d_ctx->plan(), d_data, d_numVectors));
      }
      size_type
      localSize() const
      {
        return d_ctx->localSize();
      }
      size_type
      locallyOwnedSize() const
      {
        return d_ctx->locallyOwnedSize();
      }
      size_type
      ghostSize() const
      {
        return d_ctx->localSize() - d_ctx->locallyOwnedSize();
      }
      const std::shared_ptr<const DeviceContext> &
      getDeviceContext() const
      {
        return d_ctx;
      }
      bool
      isCompatible(const MultiVector &rhs) const
      {
        return d_ctx == rhs.d_ctx;
      }
      friend void
      swap(MultiVector &X, MultiVector &Y)
      {
        std::swap(X.d_ctx, Y.d_ctx);
        std::swap(X.d_numVectors, Y.d_numVectors);
        std::swap(X.d_data, Y.d_data);
      }

    private:
      size_t
      bytes() const
      {
        return (size_t)localSize() * d_numVectors * sizeof(double);
      }
      std::shared_ptr<const DeviceContext> d_ctx;
      size_type                            d_numVectors;
      double *                             d_data = nullptr;
    };

    // src/linearAlgebra/OperatorContext.h:48-111
    template <typename ValueTypeOperator, typename ValueTypeOperand, utils::MemorySpace memorySpace>
    class OperatorContext
    {
    public:
      using ValueTypeUnion = double;
      virtual ~OperatorContext() = default;
      // X may be modified (ghost update + hanging-node fill); Y is fully overwritten
      virtual void
      apply(MultiVector<ValueTypeOperand, memorySpace> &X,
            MultiVector<ValueTypeUnion, memorySpace> &  Y,
            bool                                        updateGhostX = false,
            bool                                        updateGhostY = false) const = 0;
    };

    using DeviceMultiVector     = MultiVector<double, utils::MemorySpace::DEVICE>;
    using DeviceOperatorContext = OperatorContext<double, double, utils::MemorySpace::DEVICE>;

    // operators implemented natively by libhxb200 expose their handle so the filters can run fused
    class NativeOperator : public DeviceOperatorContext
    {
    public:
      hx_op *
      handle() const
      {
        return d_op;
      }
      void
      apply(DeviceMultiVector &X, DeviceMultiVector &Y, bool updateGhostX = false, bool updateGhostY = false) const override
      {
        utils::throwException(X.getNumberComponents() == Y.getNumberComponents(),
                              "apply: X and Y hold different numbers of vectors");
        utils::hxCheck(hx_op_apply(d_op, X.data(), Y.data(), X.getNumberComponents(), updateGhostX, updateGhostY));
      }
      ~NativeOperator() override
      {
        if (d_op)
          hx_op_destroy(d_op);
      }

    protected:
      NativeOperator() = default;
      hx_op *                              d_op = nullptr;
      std::shared_ptr<const DeviceContext> d_ctx;
    };
  } // namespace linearAlgebra

  namespace ksdft
  {
    // One summand of the cell Hamiltonian (Hamiltonian<T>::getLocal(), src/ksdft/Hamiltonian.h:33-52): S2 doubles,
    // cells concatenated, each n_c x n_c row-major, in host or device memory.
    struct LocalHamiltonianComponent
    {
      const double *cellMatrices = nullptr;
      bool          onDevice     = false;
    };

    // KohnShamOperatorContextFE<double,double,double,double,DEVICE,3>
    class KohnShamOperatorContextFE : public linearAlgebra::NativeOperator
    {
    public:
      KohnShamOperatorContextFE(std::shared_ptr<const linearAlgebra::DeviceContext> ctx,
                                const std::vector<LocalHamiltonianComponent> &      hamiltonianComponentsVec,
                                const basis::AtomCenterNonLocalArrays *             nonLocal = nullptr)
      {
        d_ctx = std::move(ctx);
        utils::hxCheck(hx_cellop_create(d_ctx->plan(), &d_op));
        if (nonLocal)
          {
            hx_nonlocal_desc nl;
            std::memset(&nl, 0, sizeof(nl));
            nl.struct_size         = sizeof(nl);
            nl.proj_halo           = nonLocal->mpiPatternP2PProj.desc();
            nl.num_cell_proj       = nonLocal->numProjsInCells.data();
            nl.cell_proj_local_ids = nonLocal->locallyOwnedCellLocalProjectorIds.data();
            nl.cell_c              = nonLocal->cellWiseC.data();
            nl.v                   = nonLocal->V.data();
            utils::hxCheck(hx_cellop_set_nonlocal(d_op, &nl));
          }
        reinit(hamiltonianComponentsVec);
      }

      // src/ksdft/KohnShamOperatorContextFE.t.cpp:1240-1311: sum the components' cell matrices, (re)tile them
      void
      reinit(const std::vector<LocalHamiltonianComponent> &comps)
      {
        utils::throwException(!comps.empty(), "reinit: no Hamiltonian component");
        const size_t S2 = d_ctx->cellMatrixSize();
        if (comps.size() == 1)
          {
            utils::hxCheck(hx_cellop_set_matrices(d_op, comps[0].cellMatrices, comps[0].onDevice ? 1 : 0));
            return;
          }
        double *sum = nullptr, *tmp = nullptr;
        utils::hxCheck(hx_device_alloc((void **)&sum, S2 * sizeof(double)));
        utils::hxCheck(hx_memset_zero(sum, S2 * sizeof(double)));
        for (const auto &c : comps)
          {
            const double *src = c.cellMatrices;
            if (!c.onDevice)
              {
                if (!tmp)
                  utils::hxCheck(hx_device_alloc((void **)&tmp, S2 * sizeof(double)));
                utils::hxCheck(hx_memcpy_h2d(tmp, c.cellMatrices, S2 * sizeof(double)));
                src = tmp;
              }
            // flat axpby over S2 entries, in slabs that fit the 32-bit row count of the ABI
            for (size_t o = 0; o < S2; o += (size_t)1 << 30)
              {
                const uint32_t n = (uint32_t)std::min<size_t>((size_t)1 << 30, S2 - o);
                utils::hxCheck(hx_axpby(d_ctx->plan(), n, 1, 1.0, sum + o, 1.0, src + o, sum + o));
              }
          }
        utils::hxCheck(hx_plan_synchronize(d_ctx->plan()));
        utils::hxCheck(hx_cellop_set_matrices(d_op, sum, 1));
        hx_device_free(sum);
        if (tmp)
          hx_device_free(tmp);
      }
    };
  } // namespace ksdft

  namespace basis
  {
    // mass-lumped (GLL) overlap / overlap-inverse operators, apply only
    class DiagonalOverlapOperator : public linearAlgebra::NativeOperator
    {
    protected:
      DiagonalOverlapOperator(std::shared_ptr<const linearAlgebra::DeviceContext> ctx,
                              const std::vector<double> &                         diagonal,
                              const std::vector<double> *                         atomBlock,
                              int                                                 variant)
      {
        d_ctx = std::move(ctx);
        utils::throwException(diagonal.size() == d_ctx->localSize(), "diagonal must cover the local (owned+ghost) rows");
        utils::hxCheck(hx_diagop_create(d_ctx->plan(), diagonal.data(), atomBlock ? atomBlock->data() : nullptr, variant, &d_op));
      }
    };
    // src/basis/CFEOverlapInverseOpContextGLL.t.cpp:499-560
    class CFEOverlapInverseOpContextGLL : public DiagonalOverlapOperator
    {
    public:
      CFEOverlapInverseOpContextGLL(std::shared_ptr<const linearAlgebra::DeviceContext> ctx,
                                    const std::vector<double> &                         diagonalInv)
        : DiagonalOverlapOperator(std::move(ctx), diagonalInv, nullptr, HX_DIAG_CFE)
      {}
    };
    // src/basis/OEFEAtomBlockOverlapInvOpContextGLL.t.cpp:953-1108
    class OEFEAtomBlockOverlapInvOpContextGLL : public DiagonalOverlapOperator
    {
    public:
      OEFEAtomBlockOverlapInvOpContextGLL(std::shared_ptr<const linearAlgebra::DeviceContext> ctx,
                                          const std::vector<double> &                         diagonalInv,
                                          const std::vector<double> &atomBlockEnrichmentOverlapInv)
        : DiagonalOverlapOperator(std::move(ctx), diagonalInv, &atomBlockEnrichmentOverlapInv, HX_DIAG_OEFE_ATOMBLOCK)
      {}
    };
    // src/basis/OrthoEFEOverlapOperatorContext.t.cpp:2085-2235 (mass-lumped branch)
    class OrthoEFEOverlapOperatorContext : public DiagonalOverlapOperator
    {
    public:
      OrthoEFEOverlapOperatorContext(std::shared_ptr<const linearAlgebra::DeviceContext> ctx,
                                     const std::vector<double> &                         diagonal,
                                     const std::vector<double> &                         atomBlockEnrichmentOverlap)
        : DiagonalOverlapOperator(std::move(ctx), diagonal, &atomBlockEnrichmentOverlap, HX_DIAG_OEFE_MASS)
      {}
    };
  } // namespace basis

  namespace linearAlgebra
  {
    namespace blasLapack
    {
      // z = alpha x + beta y over the first nRows rows (src/linearAlgebra/BlasLapackKernels.cpp:456-475)
      inline void
      axpby(const DeviceContext &ctx, size_type nRows, size_type B, double alpha, const double *x, double beta,
            const double *y, double *z)
      {
        utils::hxCheck(hx_axpby(ctx.plan(), nRows, B, alpha, x, beta, y, z));
      }
    } // namespace blasLapack

    // src/linearAlgebra/ChebyshevFilter.t.cpp:39-134.  Native operators run the fused device path; any other
    // OperatorContext subclass runs the same recurrence through its apply().
    inline void
    ChebyshevFilter(const DeviceOperatorContext &A,
                    const DeviceOperatorContext &BInv,
                    DeviceMultiVector &          eigenSubspaceGuess,
                    const size_type              polynomialDegree,
                    const double                 wantedSpectrumLowerBound,
                    const double                 wantedSpectrumUpperBound,
                    const double                 unWantedSpectrumUpperBound,
                    DeviceMultiVector &          filteredSubspace)
    {
      auto *      nA = dynamic_cast<const NativeOperator *>(&A);
      auto *      nB = dynamic_cast<const NativeOperator *>(&BInv);
      const auto  B  = eigenSubspaceGuess.getNumberComponents();
      if (nA && nB)
        {
          utils::hxCheck(hx_chebyshev_filter(nA->handle(), nB->handle(), eigenSubspaceGuess.data(), filteredSubspace.data(), B,
                                             polynomialDegree, wantedSpectrumLowerBound, wantedSpectrumUpperBound,
                                             unWantedSpectrumUpperBound));
          return;
        }
      const DeviceContext &ctx   = *eigenSubspaceGuess.getDeviceContext();
      const size_type      nOwn  = eigenSubspaceGuess.locallyOwnedSize();
      const double         e     = (unWantedSpectrumUpperBound - wantedSpectrumUpperBound) / 2.0;
      const double         c     = (unWantedSpectrumUpperBound + wantedSpectrumUpperBound) / 2.0;
      double               sigma = e / (wantedSpectrumLowerBound - c);
      const double         sigma1 = sigma, gamma = 2.0 / sigma1;
      DeviceMultiVector    scratch1(eigenSubspaceGuess.getDeviceContext(), B), scratch2(eigenSubspaceGuess.getDeviceContext(), B);
      A.apply(eigenSubspaceGuess, scratch1, true, false);
      BInv.apply(scratch1, scratch2, false, false);
      blasLapack::axpby(ctx, nOwn, B, sigma1 / e, scratch2.data(), -sigma1 / e * c, eigenSubspaceGuess.data(),
                        filteredSubspace.data());
      for (size_type degree = 2; degree <= polynomialDegree; degree++)
        {
          const double sigma2 = 1.0 / (gamma - sigma);
          A.apply(filteredSubspace, scratch1, true, false);
          BInv.apply(scratch1, scratch2, false, false);
          blasLapack::axpby(ctx, nOwn, B, 2.0 * sigma2 / e, scratch2.data(), -2.0 * sigma2 / e * c, filteredSubspace.data(),
                            scratch1.data());
          blasLapack::axpby(ctx, nOwn, B, 1.0, scratch1.data(), -sigma * sigma2, eigenSubspaceGuess.data(),
                            eigenSubspaceGuess.data());
          swap(eigenSubspaceGuess, filteredSubspace);
          sigma = sigma2;
        }
      std::vector<double> tmp((size_t)filteredSubspace.localSize() * B);
      filteredSubspace.copyTo(tmp.data());
      eigenSubspaceGuess.copyFrom(tmp.data());
    }

    // src/linearAlgebra/ChebyshevFilter.t.cpp:242-445 (native operators only)
    inline void
    ResidualChebyshevFilterGEP(const DeviceOperatorContext &A,
                               const DeviceOperatorContext &B,
                               const DeviceOperatorContext &BInv,
                               std::vector<double> &        eigenvalues,
                               DeviceMultiVector &          eigenSubspaceGuess,
                               const size_type              polynomialDegree,
                               const double                 wantedSpectrumLowerBound,
                               const double                 wantedSpectrumUpperBound,
                               const double                 unWantedSpectrumUpperBound,
                               DeviceMultiVector &          filteredSubspace)
    {
      auto *nA = dynamic_cast<const NativeOperator *>(&A);
      auto *nB = dynamic_cast<const NativeOperator *>(&B);
      auto *nI = dynamic_cast<const NativeOperator *>(&BInv);
      utils::throwException(nA && nB && nI, "ResidualChebyshevFilterGEP<DEVICE> needs libhxb200 operators");
      utils::throwException(eigenvalues.size() == eigenSubspaceGuess.getNumberComponents(), "one eigenvalue per vector");
      utils::hxCheck(hx_residual_chebyshev_filter(nA->handle(), nB->handle(), nI->handle(), eigenvalues.data(),
                                                  eigenSubspaceGuess.data(), filteredSubspace.data(),
                                                  eigenSubspaceGuess.getNumberComponents(), polynomialDegree,
                                                  wantedSpectrumLowerBound, wantedSpectrumUpperBound,
                                                  unWantedSpectrumUpperBound));
    }

    namespace RayleighRitzEigenSolverInternal
    {
      // computeXTransOpX (src/linearAlgebra/RayleighRitzEigenSolver.t.cpp:685-844): B x B column-major, lower
      // triangle filled, summed over ranks.
      inline std::vector<double>
      computeXTransOpX(DeviceMultiVector &X, const NativeOperator &Op, size_type eigenVectorBatchSize)
      {
        const size_type     B = X.getNumberComponents();
        std::vector<double> S((size_t)B * B);
        utils::hxCheck(hx_xtopx(Op.handle(), X.data(), B, eigenVectorBatchSize, S.data()));
        return S;
      }
    } // namespace RayleighRitzEigenSolverInternal

    namespace elpaScalaOpInternal
    {
      // subspaceRotation (src/linearAlgebra/ElpaScalapackOperations.t.cpp:185-335); Q: B x B column-major, replicated
      inline void
      subspaceRotation(DeviceMultiVector &X, const std::vector<double> &rotationMat, bool rotationMatTranspose,
                       bool isRotationMatLowerTria)
      {
        const size_type B = X.getNumberComponents();
        utils::throwException(rotationMat.size() == (size_t)B * B, "rotation matrix must be B x B");
        utils::hxCheck(hx_subspace_rotation(X.getDeviceContext()->plan(), X.data(), B, rotationMat.data(), rotationMatTranspose,
                                            isRotationMatLowerTria));
      }
    } // namespace elpaScalaOpInternal
  }   // namespace linearAlgebra

  // ---------------------------------------------------------------------------------------------------------------
  // electrostatics: the Poisson solve that produces the Hartree / nuclear potentials every SCF
  // iteration runs the same gather -> cell GEMM -> scatter kernel with grad N_i . grad N_j cell matrices
  // ---------------------------------------------------------------------------------------------------------------
  namespace basis
  {
    // one more ConstraintsLocal on the DoF numbering of the DeviceContext (e.g. feBasisManagerX of the Poisson
    // problem: inhomogeneous Dirichlet data); the six arrays of src/basis/CFEConstraintsLocalDealii.t.cpp:286-462
    struct ConstraintsLocalArrays
    {
      std::vector<size_type> rowConstraintsIdsLocal, rowConstraintsSizes, columnConstraintsAccumulated, columnConstraintsIdsLocal;
      std::vector<double>    columnConstraintsValues, constraintsInhomogenities;
    };
    inline uint32_t
    registerConstraints(const linearAlgebra::DeviceContext &ctx, const ConstraintsLocalArrays &c)
    {
      uint32_t id = 0;
      utils::hxCheck(hx_plan_add_constraints(ctx.plan(), (uint32_t)c.rowConstraintsIdsLocal.size(), c.rowConstraintsIdsLocal.data(),
                                             c.rowConstraintsSizes.data(), c.columnConstraintsAccumulated.data(),
                                             c.columnConstraintsIdsLocal.data(), c.columnConstraintsValues.data(),
                                             c.constraintsInhomogenities.data(), &id));
      return id;
    }
  } // namespace basis

  namespace electrostatics
  {
    // LaplaceOperatorContextFE<double,double,DEVICE,3> (src/electrostatics/LaplaceOperatorContextFE.t.cpp:395-470):
    // constraintsX / constraintsY are ids from basis::registerConstraints (0 = the DeviceContext's own constraints)
    class LaplaceOperatorContextFE : public linearAlgebra::NativeOperator
    {
    public:
      LaplaceOperatorContextFE(std::shared_ptr<const linearAlgebra::DeviceContext> ctx, const double *gradNiGradNjInAllCells,
                               bool onDevice, uint32_t constraintsX = 0, uint32_t constraintsY = 0)
      {
        d_ctx = std::move(ctx);
        utils::hxCheck(hx_cellop_create(d_ctx->plan(), &d_op));
        utils::hxCheck(hx_cellop_set_matrices(d_op, gradNiGradNjInAllCells, onDevice ? 1 : 0));
        utils::hxCheck(hx_cellop_set_constraint_sets(d_op, constraintsX, constraintsY));
      }
    };
  } // namespace electrostatics

  namespace linearAlgebra
  {
    // PreconditionerJacobi (src/linearAlgebra/PreconditionerJacobi.t.cpp:35-82): takes the DIAGONAL like the reference
    class PreconditionerJacobi : public NativeOperator
    {
    public:
      PreconditionerJacobi(std::shared_ptr<const DeviceContext> ctx, const std::vector<double> &diagonal)
      {
        d_ctx = std::move(ctx);
        utils::throwException(diagonal.size() == d_ctx->localSize(), "diagonal must cover the local (owned+ghost) rows");
        std::vector<double> inv(diagonal.size());
        for (size_t i = 0; i < inv.size(); ++i)
          inv[i] = 1.0 / diagonal[i]; // blasLapack::reciprocalX in the reference constructor (:40-45)
        utils::hxCheck(hx_diagop_create(d_ctx->plan(), inv.data(), nullptr, HX_DIAG_JACOBI, &d_op));
      }
    };

    // LinearSolverFunction (src/linearAlgebra/LinearSolverFunction.h): the handles CGLinearSolver::solve asks for
    class LinearSolverFunction
    {
    public:
      virtual ~LinearSolverFunction() = default;
      virtual const NativeOperator &
      getAxContext() const = 0;
      virtual const NativeOperator &
      getPCContext() const = 0;
      virtual const DeviceMultiVector &
      getRhs() const = 0;
      virtual const DeviceMultiVector &
      getInitialGuess() const = 0;
      virtual void
      setSolution(const DeviceMultiVector &x) = 0;
    };

    // LinearSolverErrorCode / LinearSolverError (src/linearAlgebra/LinearAlgebraTypes.h:83-90, 120-160)
    enum class LinearSolverErrorCode
    {
      SUCCESS,
      FAILED_TO_CONVERGE,
      RESIDUAL_DIVERGENCE,
      DIVISON_BY_ZERO,
      OTHER_ERROR
    };
    struct LinearSolverError
    {
      bool                  isSuccess;
      LinearSolverErrorCode err;
      std::string           msg;
    };

    // CGLinearSolver (src/linearAlgebra/CGLinearSolver.h:60-100, .t.cpp:68-300)
    class CGLinearSolver
    {
    public:
      CGLinearSolver(size_type maxIter, double absoluteTol, double relativeTol, double divergenceTol)
        : d_maxIter(maxIter)
        , d_absoluteTol(absoluteTol)
        , d_relativeTol(relativeTol)
        , d_divergenceTol(divergenceTol)
      {}
      LinearSolverError
      solve(LinearSolverFunction &f)
      {
        const DeviceMultiVector &b = f.getRhs();
        DeviceMultiVector        x(f.getInitialGuess());
        uint32_t                 iters  = 0;
        int                      status = 0;
        utils::hxCheck(hx_cg_solve(f.getAxContext().handle(), f.getPCContext().handle(), b.data(), x.data(),
                                   b.getNumberComponents(), d_maxIter, d_absoluteTol, d_relativeTol, d_divergenceTol, &iters,
                                   &status, nullptr));
        f.setSolution(x);
        d_iterations = iters;
        LinearSolverError e;
        e.err       = static_cast<LinearSolverErrorCode>(status);
        e.isSuccess = status == HX_CG_SUCCESS;
        e.msg       = e.isSuccess ? "CGLinear solve converged in maximum " + std::to_string(iters) + " iterations" :
                                    (status == HX_CG_FAILED_TO_CONVERGE ? "The linear solver failed to converge" :
                                     status == HX_CG_RESIDUAL_DIVERGENCE ? "The residual diverged" : "Other error");
        return e;
      }
      size_type
      iterations() const
      {
        return d_iterations;
      }

    private:
      size_type d_maxIter, d_iterations = 0;
      double    d_absoluteTol, d_relativeTol, d_divergenceTol;
    };
    // ---------------------------------------------------------------------------------------------------------------
    // The eigensolve around the path: same class names, argument order and error structs as
    // the reference; all block-vector work and the dense B x B steps stay on the device (hx_chfsi_solve & co).
    // ---------------------------------------------------------------------------------------------------------------
    // src/linearAlgebra/LinearAlgebraTypes.h:92-160
    enum class EigenSolverErrorCode
    {
      SUCCESS,
      LAPACK_ERROR,
      LANCZOS_BETA_ZERO,
      LANCZOS_SUBSPACE_INSUFFICIENT,
      CHFSI_ORTHONORMALIZATION_ERROR,
      CHFSI_RAYLEIGH_RITZ_ERROR,
      KS_MAX_PASS_ERROR,
      KS_CHFSI_ERROR,
      KS_LANCZOS_ERROR,
      KS_NEWTON_RAPHSON_ERROR,
      ELPASCALAPACK_ERROR,
      OTHER_ERROR
    };
    struct EigenSolverError
    {
      bool                 isSuccess;
      EigenSolverErrorCode err;
      std::string          msg;
    };
    enum class OrthonormalizationErrorCode
    {
      SUCCESS,
      LAPACK_ERROR,
      NON_ORTHONORMALIZABLE_MULTIVECTOR,
      ELPASCALAPACK_ERROR,
      MAX_PASS_EXCEEDED
    };
    struct OrthonormalizationError
    {
      bool                        isSuccess;
      OrthonormalizationErrorCode err;
      std::string                 msg;
    };
    enum class OrthogonalizationType
    {
      CHOLESKY_GRAMSCHMIDT,
      MULTIPASS_CGS
    };
    namespace EigenSolverErrorMsg
    {
      inline EigenSolverError
      isSuccessAndMsg(EigenSolverErrorCode c)
      {
        static const char *msgs[] = {"Success",
                                     "Lapack error",
                                     "Lanczos beta reached zero",
                                     "Lanczos Krylov subspace insufficient for the tolerance",
                                     "ChFSI orthonormalization error: ",
                                     "ChFSI Rayleigh-Ritz error: ",
                                     "Maximum Chebyshev filter passes reached",
                                     "ChFSI error: ",
                                     "Lanczos error: ",
                                     "Newton-Raphson error: ",
                                     "ELPA/ScaLAPACK error",
                                     "Other error"};
        EigenSolverError e;
        e.err       = c;
        e.isSuccess = c == EigenSolverErrorCode::SUCCESS; // as in the reference: only SUCCESS counts
        e.msg       = msgs[static_cast<int>(c)];
        return e;
      }
    } // namespace EigenSolverErrorMsg

    inline const NativeOperator &
    asNative(const DeviceOperatorContext &op, const char *what)
    {
      auto *n = dynamic_cast<const NativeOperator *>(&op);
      utils::throwException(n != nullptr, std::string(what) + "<DEVICE> needs libhxb200 operators");
      return *n;
    }

    // OrthonormalizationFunctions (src/linearAlgebra/OrthonormalizationFunctions.h, .t.cpp:154-352)
    class OrthonormalizationFunctions
    {
    public:
      explicit OrthonormalizationFunctions(size_type eigenVectorBatchSize)
        : d_eigenVecBatchSize(eigenVectorBatchSize)
      {}
      OrthonormalizationError
      CholeskyGramSchmidt(DeviceMultiVector &X, DeviceMultiVector &orthogonalizedX, const DeviceOperatorContext &B)
      {
        int st = 0;
        utils::hxCheck(hx_cholesky_gram_schmidt(asNative(B, "CholeskyGramSchmidt").handle(), X.data(), orthogonalizedX.data(),
                                                X.getNumberComponents(), batch(X), &st));
        OrthonormalizationError e;
        e.err       = static_cast<OrthonormalizationErrorCode>(st);
        e.isSuccess = st == 0;
        e.msg       = st == 0 ? "Success" : hx_last_error();
        return e;
      }

    private:
      size_type
      batch(const DeviceMultiVector &X) const
      {
        return d_eigenVecBatchSize ? d_eigenVecBatchSize : X.getNumberComponents();
      }
      size_type d_eigenVecBatchSize;
    };

    // RayleighRitzEigenSolver, standard problem (src/linearAlgebra/RayleighRitzEigenSolver.t.cpp:70-290)
    class RayleighRitzEigenSolver
    {
    public:
      explicit RayleighRitzEigenSolver(size_type eigenVectorBatchSize)
        : d_eigenVecBatchSize(eigenVectorBatchSize)
      {}
      EigenSolverError
      solve(const DeviceOperatorContext &A, DeviceMultiVector &X, std::vector<double> &eigenValues,
            DeviceMultiVector &eigenVectors, bool computeEigenVectors = false)
      {
        const size_type B = X.getNumberComponents();
        eigenValues.resize(B);
        int st = 0;
        utils::hxCheck(hx_rayleigh_ritz(asNative(A, "RayleighRitzEigenSolver").handle(), X.data(), eigenVectors.data(), B,
                                        d_eigenVecBatchSize ? d_eigenVecBatchSize : B, eigenValues.data(),
                                        computeEigenVectors, &st));
        return EigenSolverErrorMsg::isSuccessAndMsg(st == 0 ? EigenSolverErrorCode::SUCCESS :
                                                              EigenSolverErrorCode::LAPACK_ERROR);
      }

    private:
      size_type d_eigenVecBatchSize;
    };

    // ChebyshevFilteredEigenSolver (src/linearAlgebra/ChebyshevFilteredEigenSolver.h:84-125, .t.cpp:39-438)
    class ChebyshevFilteredEigenSolver
    {
    public:
      ChebyshevFilteredEigenSolver(const double wantedSpectrumLowerBound, const double wantedSpectrumUpperBound,
                                   const double unWantedSpectrumUpperBound, const double polynomialDegree,
                                   const double illConditionTolerance, DeviceMultiVector &eigenSubspaceGuess,
                                   bool isResidualChebyshevFilter = true, const size_type eigenVectorBatchSize = 0,
                                   OrthogonalizationType orthoType = OrthogonalizationType::CHOLESKY_GRAMSCHMIDT)
        : d_eigenVecBatchSize(eigenVectorBatchSize)
        , d_isResidualChebyFilter(isResidualChebyshevFilter)
        , d_orthoType(orthoType)
      {
        reinit(wantedSpectrumLowerBound, wantedSpectrumUpperBound, unWantedSpectrumUpperBound, polynomialDegree,
               illConditionTolerance, eigenSubspaceGuess);
      }
      void
      reinit(const double wantedSpectrumLowerBound, const double wantedSpectrumUpperBound,
             const double unWantedSpectrumUpperBound, const double polynomialDegree, const double illConditionTolerance,
             DeviceMultiVector &eigenSubspaceGuess)
      {
        d_wantedSpectrumLowerBound   = wantedSpectrumLowerBound;
        d_wantedSpectrumUpperBound   = wantedSpectrumUpperBound;
        d_unWantedSpectrumUpperBound = unWantedSpectrumUpperBound;
        d_polynomialDegree           = polynomialDegree;
        d_illConditionTolerance      = illConditionTolerance;
        d_eigenSubspaceGuess         = &eigenSubspaceGuess;
      }
      // eigenValues: in (read by the residual filter) / out (Ritz values); eigenVectors: out
      EigenSolverError
      solve(const DeviceOperatorContext &A, std::vector<double> &eigenValues, DeviceMultiVector &eigenVectors,
            bool computeEigenVectors, const DeviceOperatorContext &B, const DeviceOperatorContext &BInv)
      {
        const size_type nVec = eigenVectors.getNumberComponents();
        utils::throwException(d_eigenSubspaceGuess->getNumberComponents() == nVec, "guess and eigenVectors differ in width");
        eigenValues.resize(nVec, 0.0);
        int st = 0;
        utils::hxCheck(hx_chfsi_solve_ortho(asNative(A, "ChebyshevFilteredEigenSolver").handle(),
                                      asNative(B, "ChebyshevFilteredEigenSolver").handle(),
                                      asNative(BInv, "ChebyshevFilteredEigenSolver").handle(), d_eigenSubspaceGuess->data(),
                                      eigenVectors.data(), nVec, d_eigenVecBatchSize ? d_eigenVecBatchSize : nVec,
                                      (size_type)d_polynomialDegree, d_wantedSpectrumLowerBound, d_wantedSpectrumUpperBound,
                                      d_unWantedSpectrumUpperBound, d_isResidualChebyFilter, eigenValues.data(),
                                      computeEigenVectors,
                                      d_orthoType == OrthogonalizationType::MULTIPASS_CGS ? HX_ORTHO_MULTIPASS_CGS :
                                                                                            HX_ORTHO_CHOLESKY_GRAMSCHMIDT,
                                      &st));
        EigenSolverError e = EigenSolverErrorMsg::isSuccessAndMsg(static_cast<EigenSolverErrorCode>(st));
        if (st != 0)
          e.msg += hx_last_error();
        return e;
      }

    private:
      double                d_wantedSpectrumLowerBound, d_wantedSpectrumUpperBound, d_unWantedSpectrumUpperBound;
      double                d_polynomialDegree, d_illConditionTolerance;
      DeviceMultiVector *   d_eigenSubspaceGuess = nullptr;
      const size_type       d_eigenVecBatchSize;
      bool                  d_isResidualChebyFilter;
      OrthogonalizationType d_orthoType;
    };

    // LanczosExtremeEigenSolver (src/linearAlgebra/LanczosExtremeEigenSolver.h:60-150, .t.cpp:216-520); eigenvalues
    // only (computeEigenVectors is what no reference driver asks of it: KohnShamEigenSolver.t.cpp:255-260)
    class LanczosExtremeEigenSolver
    {
    public:
      LanczosExtremeEigenSolver(const size_type maxKrylovSubspaceSize, const size_type numLowerExtermeEigenValues,
                                const size_type numUpperExtermeEigenValues, std::vector<double> &tolerance,
                                double lanczosBetaTolerance, const DeviceMultiVector &initialGuess,
                                bool isAdaptiveSolve = true)
        : d_initialGuess(initialGuess)
        , d_isAdaptiveSolve(isAdaptiveSolve)
      {
        utils::throwException(initialGuess.getNumberComponents() == 1, "the Lanczos guess is one vector");
        d_maxKrylovSubspaceSize      = maxKrylovSubspaceSize;
        d_numLowerExtermeEigenValues = numLowerExtermeEigenValues;
        d_numUpperExtermeEigenValues = numUpperExtermeEigenValues;
        d_tolerance                  = tolerance;
        d_lanczosBetaTolerance       = lanczosBetaTolerance;
      }
      EigenSolverError
      solve(const DeviceOperatorContext &A, std::vector<double> &eigenValues, DeviceMultiVector & /*eigenVectors*/,
            bool computeEigenVectors, const DeviceOperatorContext &B, const DeviceOperatorContext &BInv)
      {
        utils::throwException(!computeEigenVectors, "LanczosExtremeEigenSolver<DEVICE>: eigenvalues only");
        const size_type nW = d_numLowerExtermeEigenValues + d_numUpperExtermeEigenValues;
        utils::throwException(d_maxKrylovSubspaceSize >= nW,
                              "Maximum Krylov subspace size should be more than number of required eigenPairs.");
        eigenValues.assign(nW, 0.0);
        d_diagonal.assign(d_maxKrylovSubspaceSize, 0.0);
        d_subDiagonal.assign(d_maxKrylovSubspaceSize, 0.0);
        uint32_t k  = 0;
        int      st = 0;
        utils::hxCheck(hx_lanczos_extreme(asNative(A, "Lanczos").handle(), asNative(B, "Lanczos").handle(),
                                          asNative(BInv, "Lanczos").handle(), d_initialGuess.data(), d_maxKrylovSubspaceSize,
                                          d_numLowerExtermeEigenValues, d_numUpperExtermeEigenValues, d_tolerance.data(),
                                          d_lanczosBetaTolerance, d_isAdaptiveSolve, eigenValues.data(), d_diagonal.data(),
                                          d_subDiagonal.data(), &k, &st));
        d_diagonal.resize(k);
        d_subDiagonal.resize(k);
        d_isSolved = true;
        return EigenSolverErrorMsg::isSuccessAndMsg(static_cast<EigenSolverErrorCode>(st));
      }
      void
      getTridiagonalMatrix(std::vector<double> &diagonal, std::vector<double> &subDiagonal) const
      {
        utils::throwException(d_isSolved, "Cannot call getTridiagonalMatrix() before solving the eigenproblem.");
        diagonal    = d_diagonal;
        subDiagonal = d_subDiagonal;
      }

    private:
      DeviceMultiVector   d_initialGuess;
      size_type           d_maxKrylovSubspaceSize, d_numLowerExtermeEigenValues, d_numUpperExtermeEigenValues;
      std::vector<double> d_tolerance, d_diagonal, d_subDiagonal;
      double              d_lanczosBetaTolerance;
      bool                d_isAdaptiveSolve, d_isSolved = false;
    };
  } // namespace linearAlgebra

  namespace basis
  {
    // The arrays FEBasisDataStorage hands to FEBasisOperations (src/basis/FEBasisOperations.t.cpp:111-160): quadrature
    // points per cell, JxW, basis values per cell (nq x n_c, DoF index fastest; ONE matrix when the same quadrature rule
    // and DoF count hold in every cell)
    struct FEBasisDataStorageArrays
    {
      std::vector<size_type> nCellQuadraturePoints;
      std::vector<double>    basisDataInAllCells, JxWInAllCells;
      bool                   sameBasisDataInAllCells = false;
    };

    // FEBasisOperations (src/basis/FEBasisOperations.h): the two members on either side of the H.X path
    class FEBasisOperations
    {
    public:
      FEBasisOperations(std::shared_ptr<const linearAlgebra::DeviceContext> ctx, const FEBasisDataStorageArrays &st)
        : d_ctx(std::move(ctx))
      {
        utils::throwException(st.nCellQuadraturePoints.size() > 0, "no cells");
        hx_fe_basis_desc d;
        std::memset(&d, 0, sizeof(d));
        d.struct_size             = sizeof(d);
        d.same_basis_in_all_cells = st.sameBasisDataInAllCells;
        d.num_cell_quad           = st.nCellQuadraturePoints.data();
        d.basis_data              = st.basisDataInAllCells.data();
        d.jxw                     = st.JxWInAllCells.data();
        utils::hxCheck(hx_fe_basis_create(d_ctx->plan(), &d, &d_basis));
        d_nQuad = st.JxWInAllCells.size();
      }
      FEBasisOperations(const FEBasisOperations &) = delete;
      ~FEBasisOperations()
      {
        hx_fe_basis_destroy(d_basis);
      }
      // computeFEMatrices(IDENTITY, MULT, MULT, IDENTITY, f, cellWiseFEData) (src/basis/FEBasisOperations.t.cpp:
      // 2210-2243): f = one value per quadrature point (host); cellWiseFEData = DEVICE array of cellMatrixSize() doubles;
      // addTo (device, may be null) is added cell matrix by cell matrix (the reinit component sum)
      void
      computeFEMatrices(const std::vector<double> &f, double *cellWiseFEDataDevice, const double *addToDevice = nullptr) const
      {
        utils::throwException(f.size() == d_nQuad, "one value of f per quadrature point");
        utils::hxCheck(hx_compute_fe_matrices(d_basis, f.data(), 0, addToDevice, cellWiseFEDataDevice));
      }
      hx_fe_basis *
      handle() const
      {
        return d_basis;
      }
      size_t
      nQuadraturePoints() const
      {
        return d_nQuad;
      }

    private:
      std::shared_ptr<const linearAlgebra::DeviceContext> d_ctx;
      hx_fe_basis *                                       d_basis = nullptr;
      size_t                                              d_nQuad = 0;
    };
  } // namespace basis

  namespace ksdft
  {
    // KohnShamOperatorContextFE::reinit for a local potential given at the quadrature points: computeFEMatrices, the sum
    // with the other components (addToDevice: flat cell matrices on the device, e.g. the kinetic part; may be null) and
    // the re-tiling for the cell kernel in ONE kernel (src/ksdft/KohnShamOperatorContextFE.t.cpp:1240-1311 +
    // src/basis/FEBasisOperations.t.cpp:2210-2243); the operator's nonlocal part is kept
    inline void
    reinitFromPotential(const linearAlgebra::NativeOperator &H, const basis::FEBasisOperations &feOp,
                        const std::vector<double> &fAtQuadPoints, const double *addToDevice = nullptr)
    {
      utils::throwException(fAtQuadPoints.size() == feOp.nQuadraturePoints(), "one value of f per quadrature point");
      utils::hxCheck(hx_cellop_assemble_matrices(H.handle(), feOp.handle(), fAtQuadPoints.data(), 0, addToDevice));
    }

    // DensityCalculator::computeRho (src/ksdft/DensityCalculator.h, .t.cpp:283-437)
    class DensityCalculator
    {
    public:
      explicit DensityCalculator(std::shared_ptr<const basis::FEBasisOperations> feBasisOp)
        : d_feBasisOp(std::move(feBasisOp))
      {}
      void
      computeRho(const std::vector<double> &occupation, const linearAlgebra::DeviceMultiVector &waveFunc,
                 std::vector<double> &rho) const
      {
        utils::throwException(occupation.size() == waveFunc.getNumberComponents(), "one occupation per wavefunction");
        rho.resize(d_feBasisOp->nQuadraturePoints());
        utils::hxCheck(hx_compute_rho(d_feBasisOp->handle(), waveFunc.data(), waveFunc.getNumberComponents(), occupation.data(),
                                      rho.data(), 0));
      }

    private:
      std::shared_ptr<const basis::FEBasisOperations> d_feBasisOp;
    };
  } // namespace ksdft

  namespace ksdft
  {
    // src/ksdft/Defaults.cpp:44-69
    struct LinearEigenSolverDefaults
    {
      static constexpr double    ILL_COND_TOL                 = 1e-14;
      static constexpr double    LANCZOS_EXTREME_EIGENVAL_TOL = 1e-6;
      static constexpr double    LANCZOS_BETA_TOL             = 1e-14;
      static constexpr size_type LANCZOS_MAX_KRYLOV_SUBSPACE  = 20;
    };
    struct Constants
    {
      static constexpr double BOLTZMANN_CONST_HARTREE = 3.166811429e-06;
    };
    // getChebyPolynomialDegree + CHEBY_ORDER_LOOKUP (src/ksdft/KohnShamEigenSolver.t.cpp:37-46, Defaults.cpp:51-58)
    inline size_type
    getChebyPolynomialDegree(size_type unWantedSpectrumUpperBound)
    {
      uint32_t d = 0;
      utils::hxCheck(hx_chebyshev_polynomial_degree((double)unWantedSpectrumUpperBound, &d));
      return d;
    }
    // src/ksdft/FractionalOccupancyFunction.cpp:11-40
    inline double
    fermiDirac(const double eigenValue, const double fermiEnergy, const double kb, const double T)
    {
      const double factor = (eigenValue - fermiEnergy) / (kb * T);
      return (factor >= 0) ? std::exp(-factor) / (1.0 + std::exp(-factor)) : 1.0 / (1.0 + std::exp(factor));
    }
    inline double
    fermiDiracDer(const double eigenValue, const double fermiEnergy, const double kb, const double T)
    {
      const double factor = (eigenValue - fermiEnergy) / (kb * T);
      const double beta   = 1.0 / (kb * T);
      return (factor >= 0) ? (beta * std::exp(-factor) / (1.0 + std::exp(-factor)) / (1.0 + std::exp(-factor))) :
                             (beta * std::exp(factor) / (1.0 + std::exp(factor)) / (1.0 + std::exp(factor)));
    }

    // KohnShamEigenSolver (src/ksdft/KohnShamEigenSolver.h:85-200, .t.cpp:52-560): Lanczos bounds -> Chebyshev degree
    // -> ChFSI passes until every level with a non-negligible occupancy has a converged residual.
    class KohnShamEigenSolver
    {
    public:
      using OpContext = linearAlgebra::DeviceOperatorContext;
      KohnShamEigenSolver(const size_type numElectrons, const double smearingTemperature, const double fermiEnergyTolerance,
                          const double fracOccupancyTolerance, const double eigenSolveResidualTolerance,
                          const size_type maxChebyshevFilterPass, linearAlgebra::DeviceMultiVector &waveFunctionSubspaceGuess,
                          linearAlgebra::DeviceMultiVector &lanczosGuess, bool isResidualChebyshevFilter,
                          const size_type waveFunctionBatchSize, const OpContext &MLanczos, const OpContext &MInvLanczos)
        : d_numWantedEigenvalues(waveFunctionSubspaceGuess.getNumberComponents())
        , d_eigenSolveResidualTolerance(eigenSolveResidualTolerance)
        , d_maxChebyshevFilterPass(maxChebyshevFilterPass)
        , d_waveFunctionBatchSize(waveFunctionBatchSize ? waveFunctionBatchSize : waveFunctionSubspaceGuess.getNumberComponents())
        , d_fermiEnergyTolerance(fermiEnergyTolerance)
        , d_fracOccupancyTolerance(fracOccupancyTolerance)
        , d_smearingTemperature(smearingTemperature)
        , d_fracOccupancy(d_numWantedEigenvalues)
        , d_eigSolveResNorm(d_numWantedEigenvalues)
        , d_numElectrons(numElectrons)
        , d_isResidualChebyFilter(isResidualChebyshevFilter)
        , d_waveFunctionSubspaceGuess(&waveFunctionSubspaceGuess)
        , d_lanczosGuess(&lanczosGuess)
        , d_MLanczos(&MLanczos)
        , d_MInvLanczos(&MInvLanczos)
      {}
      void
      reinitBounds(double wantedSpectrumLowerBound, double wantedSpectrumUpperBound)
      {
        d_isBoundKnown             = true;
        d_wantedSpectrumLowerBound = wantedSpectrumLowerBound;
        d_wantedSpectrumUpperBound = wantedSpectrumUpperBound;
      }
      void
      setChebyshevPolynomialDegree(size_type chebyPolyDeg)
      {
        d_setChebyPolDegExternally  = true;
        d_chebyshevPolynomialDegree = chebyPolyDeg;
      }
      void
      setResidualChebyshevFilterFlag(bool flag)
      {
        d_isResidualChebyFilter = flag;
      }

      linearAlgebra::EigenSolverError
      solve(const OpContext &kohnShamOperator, std::vector<double> &kohnShamEnergies,
            linearAlgebra::DeviceMultiVector &kohnShamWaveFunctions, bool computeWaveFunctions, const OpContext &M,
            const OpContext &MInv)
      {
        using namespace linearAlgebra;
        d_isSolved = true;
        const NativeOperator &Hn = asNative(kohnShamOperator, "KohnShamEigenSolver");
        const NativeOperator &Mn = asNative(M, "KohnShamEigenSolver");
        const size_type       B  = d_numWantedEigenvalues;
        kohnShamEnergies.resize(B, 0.0);
        std::vector<double>       tol{LinearEigenSolverDefaults::LANCZOS_EXTREME_EIGENVAL_TOL,
                                LinearEigenSolverDefaults::LANCZOS_EXTREME_EIGENVAL_TOL};
        LanczosExtremeEigenSolver lanczos(LinearEigenSolverDefaults::LANCZOS_MAX_KRYLOV_SUBSPACE, 1, 1, tol,
                                          LinearEigenSolverDefaults::LANCZOS_BETA_TOL, *d_lanczosGuess, false);
        std::vector<double>       eigenValuesLanczos(2);
        EigenSolverError          lanczosErr =
          lanczos.solve(kohnShamOperator, eigenValuesLanczos, kohnShamWaveFunctions, false, *d_MLanczos, *d_MInvLanczos);
        if (!(lanczosErr.isSuccess || lanczosErr.err == EigenSolverErrorCode::LANCZOS_SUBSPACE_INSUFFICIENT))
          {
            EigenSolverError r = EigenSolverErrorMsg::isSuccessAndMsg(EigenSolverErrorCode::KS_LANCZOS_ERROR);
            r.msg += lanczosErr.msg;
            return r;
          }
        std::vector<double> diagonal, subDiagonal;
        lanczos.getTridiagonalMatrix(diagonal, subDiagonal);
        const double residual = subDiagonal[subDiagonal.size() - 1] / 10;
        d_unWantedSpectrumUpperBound = eigenValuesLanczos[1] + residual;
        if (!d_isBoundKnown)
          {
            const double globalSize    = (double)d_globalSize(kohnShamWaveFunctions);
            d_wantedSpectrumLowerBound = eigenValuesLanczos[0];
            d_wantedSpectrumUpperBound =
              (d_unWantedSpectrumUpperBound - eigenValuesLanczos[0]) * ((double)(B * 200.0) / globalSize) + eigenValuesLanczos[0];
            if (d_wantedSpectrumUpperBound >= d_unWantedSpectrumUpperBound)
              d_wantedSpectrumUpperBound = (d_unWantedSpectrumUpperBound + eigenValuesLanczos[0]) * 0.5;
          }
        if (!d_setChebyPolDegExternally)
          {
            d_chebyshevPolynomialDegree = getChebyPolynomialDegree((size_type)d_unWantedSpectrumUpperBound);
            d_chebyshevPolynomialDegree = (size_type)(d_chebyshevPolynomialDegree * d_chebyPolyScalingFactor); // :305-306
          }
        ChebyshevFilteredEigenSolver chfsi(d_wantedSpectrumLowerBound, d_wantedSpectrumUpperBound,
                                           d_unWantedSpectrumUpperBound, (double)d_chebyshevPolynomialDegree,
                                           LinearEigenSolverDefaults::ILL_COND_TOL, *d_waveFunctionSubspaceGuess,
                                           d_isResidualChebyFilter, d_waveFunctionBatchSize);
        EigenSolverError chfsiErr = EigenSolverErrorMsg::isSuccessAndMsg(EigenSolverErrorCode::OTHER_ERROR);
        bool             nrOk     = true;
        size_type        iPass    = 0;
        for (; iPass < d_maxChebyshevFilterPass; iPass++)
          {
            chfsiErr = chfsi.solve(kohnShamOperator, kohnShamEnergies, kohnShamWaveFunctions, computeWaveFunctions, M, MInv);
            if (!chfsiErr.isSuccess)
              break;
            kohnShamWaveFunctions.updateGhostValues();
            // chemical potential by Newton-Raphson on sum 2 f(eps_i) - N_e (NewtonRaphsonSolver.t.cpp:48-96)
            nrOk = solveFermiEnergy(kohnShamEnergies);
            size_type numLevelsBelowFermiEnergy = 0, numLevelsBelowFermiEnergyResidualConverged = 0;
            for (size_type i = 0; i < B; i++)
              {
                d_fracOccupancy[i] =
                  fermiDirac(kohnShamEnergies[i], d_fermiEnergy, Constants::BOLTZMANN_CONST_HARTREE, d_smearingTemperature);
                if (d_fracOccupancy[i] > d_fracOccupancyTolerance)
                  numLevelsBelowFermiEnergy += 1;
              }
            if (computeWaveFunctions)
              {
                utils::hxCheck(hx_eigen_residual_norms(Hn.handle(), Mn.handle(), kohnShamWaveFunctions.data(), B,
                                                       d_waveFunctionBatchSize, kohnShamEnergies.data(), d_eigSolveResNorm.data()));
                for (size_type i = 0; i < B; i++)
                  if (d_fracOccupancy[i] > d_fracOccupancyTolerance && d_eigSolveResNorm[i] <= d_eigenSolveResidualTolerance)
                    numLevelsBelowFermiEnergyResidualConverged += 1;
              }
            // *d_waveFunctionSubspaceGuess = kohnShamWaveFunctions (hx_chfsi_solve leaves the Ritz vectors in the guess)
            if (numLevelsBelowFermiEnergy == numLevelsBelowFermiEnergyResidualConverged || !nrOk)
              break;
            d_wantedSpectrumLowerBound = kohnShamEnergies[0];
            d_wantedSpectrumUpperBound = kohnShamEnergies[B - 1];
            chfsi.reinit(d_wantedSpectrumLowerBound, d_wantedSpectrumUpperBound, d_unWantedSpectrumUpperBound,
                         (double)d_chebyshevPolynomialDegree, LinearEigenSolverDefaults::ILL_COND_TOL,
                         *d_waveFunctionSubspaceGuess);
          }
        d_numPasses = iPass < d_maxChebyshevFilterPass ? iPass + 1 : iPass;
        EigenSolverError r;
        if (!chfsiErr.isSuccess)
          {
            r = EigenSolverErrorMsg::isSuccessAndMsg(EigenSolverErrorCode::KS_CHFSI_ERROR);
            r.msg += chfsiErr.msg;
          }
        else if (!nrOk)
          r = EigenSolverErrorMsg::isSuccessAndMsg(EigenSolverErrorCode::KS_NEWTON_RAPHSON_ERROR);
        else if (iPass >= d_maxChebyshevFilterPass)
          r = EigenSolverErrorMsg::isSuccessAndMsg(EigenSolverErrorCode::KS_MAX_PASS_ERROR);
        else
          {
            r = EigenSolverErrorMsg::isSuccessAndMsg(EigenSolverErrorCode::SUCCESS);
            r.msg += "Number of CHFSI passes required are " + std::to_string(iPass + 1) + ".";
          }
        d_chebyPolyScalingFactor = 1.0; // :533
        return r;
      }
      double
      getFermiEnergy() const
      {
        utils::throwException(d_isSolved, "Cannot call getFermiEnergy() before solving the eigenproblem.");
        return d_fermiEnergy;
      }
      std::vector<double>
      getFractionalOccupancy() const
      {
        utils::throwException(d_isSolved, "Cannot call getFractionalOccupancy() before solving the eigenproblem.");
        return d_fracOccupancy;
      }
      std::vector<double>
      getEigenSolveResidualNorm() const
      {
        utils::throwException(d_isSolved, "Cannot call getEigenSolveResidualNorm() before solving the eigenproblem.");
        return d_eigSolveResNorm;
      }
      size_type
      getChebyshevPolynomialDegree() const
      {
        return d_chebyshevPolynomialDegree;
      }
      size_type
      getNumberOfPasses() const
      {
        return d_numPasses;
      }
      // KohnShamEigenSolver::setChebyPolyScalingFactor (src/ksdft/KohnShamEigenSolver.t.cpp:178-186): the degree from the
      // lookup table is multiplied by it in the next solve (1.34 on the first pseudopotential SCF step), then it resets to 1
      void
      setChebyPolyScalingFactor(double scalingFactor)
      {
        d_chebyPolyScalingFactor = scalingFactor;
      }
      // the global number of DoFs is MultiVector::globalSize() (all-reduced over the plan's communicator); an override is
      // kept for callers that know it
      void
      setGlobalSize(size_t n)
      {
        d_globalSizeOverride = n;
      }

    private:
      size_t
      d_globalSize(const linearAlgebra::DeviceMultiVector &X) const
      {
        if (d_globalSizeOverride)
          return d_globalSizeOverride;
        uint64_t n = 0;
        utils::hxCheck(hx_plan_global_size(X.getDeviceContext()->plan(), &n));
        return (size_t)n;
      }
      bool
      solveFermiEnergy(const std::vector<double> &eps)
      {
        const double kb = Constants::BOLTZMANN_CONST_HARTREE, T = d_smearingTemperature;
        double       x = eps[(size_t)std::ceil(static_cast<double>(d_numElectrons) / 2.0) - 1];
        for (size_t iter = 0; iter <= 20000000; ++iter)
          {
            double val = 0, force = 0;
            for (double e : eps)
              {
                val += 2 * fermiDirac(e, x, kb, T);
                force += 2 * fermiDiracDer(e, x, kb, T);
              }
            val -= (double)d_numElectrons;
            if (force == 0.0)
              return false;
            const double x1 = x - val / force;
            if (std::fabs(x1 - x) < d_fermiEnergyTolerance)
              {
                d_fermiEnergy = x1;
                return true;
              }
            x = x1;
          }
        return false;
      }
      size_type           d_numWantedEigenvalues;
      double              d_eigenSolveResidualTolerance;
      size_type           d_maxChebyshevFilterPass, d_waveFunctionBatchSize;
      double              d_fermiEnergyTolerance, d_fracOccupancyTolerance, d_smearingTemperature;
      std::vector<double> d_fracOccupancy, d_eigSolveResNorm;
      size_type           d_numElectrons;
      bool                d_isResidualChebyFilter, d_setChebyPolDegExternally = false, d_isBoundKnown = false, d_isSolved = false;
      size_type           d_chebyshevPolynomialDegree = 0, d_numPasses = 0;
      double              d_wantedSpectrumLowerBound = 0, d_wantedSpectrumUpperBound = 0, d_unWantedSpectrumUpperBound = 0;
      double              d_fermiEnergy = 0;
      size_t              d_globalSizeOverride = 0;
      double              d_chebyPolyScalingFactor = 1.0;
      linearAlgebra::DeviceMultiVector *d_waveFunctionSubspaceGuess, *d_lanczosGuess;
      const OpContext *                 d_MLanczos, *d_MInvLanczos;
    };
  } // namespace ksdft
} // namespace dftefe
#endif
