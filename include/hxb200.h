/*
 * hxb200.h — C ABI of the B200-native H.X / Chebyshev-filter / Rayleigh-Ritz hot path of dft-efe.
 *
 * One hx_plan per GPU (= per reference MPI rank).  All block vectors are DEVICE pointers to FP64
 * data in the reference MultiVector layout data[iDof*B + iVec], owned rows [0,n_owned) (range 0 =
 * classical, range 1 = enrichment) followed by ghost rows in ascending global id
 * (reference src/linearAlgebra/MultiVector.h:134-160, src/utils/MPIPatternP2P.h:107-116).
 * Index arrays in descriptors are HOST pointers (uint32_t = dftefe::size_type,
 * src/utils/TypeConfig.h:8-9) and are copied at creation.  Every call returns 0 or a negative
 * hx_status; hx_last_error() gives the message (no exceptions cross this boundary — the C++ mirror
 * in dft_efe_b200/include/ turns them into utils::throwException-style exceptions,
 * reference src/utils/Exceptions.h:128).  Calls are asynchronous on the plan's stream unless they
 * return host data.  A plan is not thread-safe (the reference operator is not re-entrant either:
 * const apply() with mutable scratch, src/ksdft/KohnShamOperatorContextFE.h:130-157).
 *
 * There is no CPU fallback: every entry point fails with HX_ERR_CUDA when no device is usable.
 *
 * Each entry point cites the reference interface it replaces.
 */
#ifndef HXB200_H
#define HXB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hx_plan hx_plan;
typedef struct hx_op   hx_op;

typedef enum hx_status {
  HX_OK              = 0,
  HX_ERR_INVALID     = -1, /* bad argument / inconsistent descriptor */
  HX_ERR_CUDA        = -2, /* CUDA runtime error, or no usable device */
  HX_ERR_COMM        = -3, /* NCCL error / communicator missing for nranks > 1 */
  HX_ERR_UNSUPPORTED = -4, /* e.g. unresolved constraint chains */
  HX_ERR_NOMEM       = -5
} hx_status;

/* The getters of utils::mpi::MPIPatternP2P the ghost communicator consumes
 * (reference src/utils/MPIPatternP2P.h:425-459). */
typedef struct hx_halo_desc {
  uint32_t        n_owned;                     /* localOwnedSize()                       */
  uint32_t        n_ghost;                     /* localGhostSize()                       */
  uint32_t        n_ghost_procs;               /* getGhostProcIds().size()               */
  const uint32_t *ghost_proc_ids;              /* getGhostProcIds()                      */
  const uint32_t *ghost_ranges;                /* getGhostLocalIndicesRanges()  [2*nGP]  */
  const uint32_t *ghost_local_ids;             /* getGhostLocalIndicesForGhostProcs()    */
  uint32_t        n_target_procs;              /* getTargetProcIds().size()              */
  const uint32_t *target_proc_ids;             /* getTargetProcIds()                     */
  const uint32_t *num_owned_for_target;        /* getNumOwnedIndicesForTargetProcs()     */
  const uint32_t *owned_local_ids_for_targets; /* getOwnedLocalIndicesForTargetProcs()   */
} hx_halo_desc;

/* The arrays basis::FEBasisManager and basis::ConstraintsLocal expose
 * (reference src/basis/FEBasisManager.h:143-160, FEBasisManager.t.cpp:414-428,545-557;
 *  src/basis/CFEConstraintsLocalDealii.t.cpp:286-462). */
typedef struct hx_mesh_desc {
  uint32_t        struct_size; /* sizeof(hx_mesh_desc), ABI guard */
  int32_t         rank, nranks;
  hx_halo_desc    halo;
  uint32_t        n_owned_classical; /* size of owned range 0; enrichment rows follow */
  uint32_t        n_cells;           /* nLocallyOwnedCells()                  */
  const uint32_t *num_cell_dofs;     /* nLocallyOwnedCellDofs(c)     [C]      */
  const uint32_t *cell_local_ids;    /* locallyOwnedCellLocalDofIds  [S]      */
  uint32_t        n_constraint_rows;
  const uint32_t *row_ids;     /* rowConstraintsIdsLocal        [nR]  */
  const uint32_t *row_sizes;   /* rowConstraintsSizes           [nR]  */
  const uint32_t *row_offsets; /* columnConstraintsAccumulated  [nR]  */
  const uint32_t *col_ids;     /* columnConstraintsIdsLocal     [nnz] */
  const double *  col_vals;    /* columnConstraintsValues       [nnz] */
  const double *  inhom;       /* constraintsInhomogenities     [nR]  */
  uint32_t        max_block;   /* largest number of vectors any call will pass (d_maxWaveFnBatch) */
} hx_mesh_desc;

/* Nonlocal pseudopotential projector data of basis::AtomCenterNonLocalOpContextFE
 * (reference src/basis/AtomCenterNonLocalOpContextFE.t.cpp:478-494,595-619). */
typedef struct hx_nonlocal_desc {
  uint32_t        struct_size;
  hx_halo_desc    proj_halo;           /* d_mpiPatternP2PProj                                   */
  const uint32_t *num_cell_proj;       /* d_numProjsInCells                  [C]                */
  const uint32_t *cell_proj_local_ids; /* d_locallyOwnedCellLocalProjectorIds [sum nProj_c]     */
  const double *  cell_c;              /* d_cellWiseC: per cell column-major nProj_c x n_c      */
  const double *  v;                   /* d_V                                [nProjLocal]       */
} hx_nonlocal_desc;

typedef enum hx_diag_variant {
  HX_DIAG_CFE            = 0, /* CFEOverlapInverseOpContextGLL::apply  (src/basis/CFEOverlapInverseOpContextGLL.t.cpp:529-558) */
  HX_DIAG_OEFE_ATOMBLOCK = 1, /* OEFEAtomBlockOverlapInvOpContextGLL::apply (src/basis/OEFEAtomBlockOverlapInvOpContextGLL.t.cpp:953-1108) */
  HX_DIAG_OEFE_MASS      = 2, /* OrthoEFEOverlapOperatorContext::apply, mass-lumped atom-block branch: as ATOMBLOCK
                                 but both ghost flags are forced to false (src/basis/OrthoEFEOverlapOperatorContext.t.cpp:2093-2235) */
  HX_DIAG_OEFE_GLOBAL    = 4, /* OrthoEFEOverlapInverseOpContextGLL::apply (src/basis/OrthoEFEOverlapInverseOpContextGLL.t.cpp:1182-1282):
                               diagonal + ONE dense block over all enrichment functions of the system; created with
                               hx_diagop_create_global_enrichment */
  HX_DIAG_JACOBI         = 3  /* linearAlgebra::PreconditionerJacobi::apply: Y = diag .* X over local rows, ghost flags
                                 honoured, no constraints (src/linearAlgebra/PreconditionerJacobi.t.cpp:52-82); pass the
                                 RECIPROCAL diagonal (the reference inverts it in the constructor, :40-45) */
} hx_diag_variant;

/* linearAlgebra::LinearSolverErrorCode of CGLinearSolver (src/linearAlgebra/LinearAlgebraTypes.h) */
typedef enum hx_cg_status {
  HX_CG_SUCCESS             = 0,
  HX_CG_FAILED_TO_CONVERGE  = 1,
  HX_CG_RESIDUAL_DIVERGENCE = 2,
  HX_CG_DIVISION_BY_ZERO    = 3,
  HX_CG_OTHER_ERROR         = 4
} hx_cg_status;

const char *hx_last_error(void);
int         hx_version(void);

/* ---- device helpers (so a host-only caller needs no CUDA runtime of its own) ---- */
int hx_device_count(int *n);
int hx_set_device(int device);
int hx_device_alloc(void **ptr, size_t bytes);
int hx_device_free(void *ptr);
int hx_host_alloc_pinned(void **ptr, size_t bytes);
int hx_host_free_pinned(void *ptr);
int hx_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes);
int hx_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes);
int hx_memset_zero(void *dst_dev, size_t bytes);

/* ---- plan ---- */
/* Replaces: the construction-time flattening in FEBasisManager (src/basis/FEBasisManager.t.cpp:161-285)
 * + MultiVector's MPICommunicatorP2P buffers (src/utils/MPICommunicatorP2P.t.cpp:39-75).
 * `stream` is a cudaStream_t (NULL = a stream the plan creates). */
int hx_plan_create(hx_plan **plan, const hx_mesh_desc *mesh, void *stream);
int hx_plan_destroy(hx_plan *plan);
int hx_plan_synchronize(hx_plan *plan);
/* Multi-GPU: rank 0 calls hx_comm_unique_id, the caller broadcasts the 128 bytes (MPI_Bcast /
 * torch.distributed), every rank calls hx_plan_attach_comm.  Replaces the MPI communicator held by
 * MPIPatternP2P (src/utils/MPIPatternP2P.h:489).  Ranks of the halo descriptors are communicator ranks. */
int hx_comm_unique_id(char id[128]);
int hx_plan_attach_comm(hx_plan *plan, const char id[128]);
/* MultiVector::globalSize() (src/linearAlgebra/MultiVector.h): locally owned rows summed over the ranks of the plan's
 * communicator (collective on first use, cached).  KohnShamEigenSolver scales wantedSpectrumUpperBound with it
 * (src/ksdft/KohnShamEigenSolver.t.cpp:270-283). */
int hx_plan_global_size(hx_plan *plan, uint64_t *n);
/* Transport the ghost communicator settled on at its first exchange: 0 = none yet / single rank, 1 = NCCL send/recv,
 * 2 = NVLink peer memory (the pack kernel stores straight into the neighbour's receive buffer and raises a flag; the
 * unpack / ordered-add kernel waits for it).  HXB200_HALO=nccl forces 1; 2 needs CUDA IPC between the ranks. */
int hx_plan_halo_transport(hx_plan *plan, int *transport);
/* Scatter strategy of the cell kernel: 0 (default) = one persistent launch, cells in the caller's order, each
 * row accumulated in ascending cell order behind per-cell completion stamps (no atomics, bitwise reproducible,
 * same per-row summation order as the reference's CPU loop, src/basis/FECellWiseDataOperations.t.cpp:87-153);
 * 1 = one launch per colour of a greedy cell colouring (cells of a launch share no row). */
int hx_plan_set_scatter_mode(hx_plan *plan, int mode);
/* Introspection used by the bit-exact parity tests of the integer work. */
int hx_plan_num_colours(hx_plan *plan, uint32_t *n);
int hx_plan_get_cell_colours(hx_plan *plan, uint32_t *colour /*[C]*/);
/* ordered scatter: the processing order (position -> cell): a delay-D list schedule of the caller's order (a
 * cell is placed only once all its placed neighbours sit >= D positions back; D = 4 x SM count, or
 * HXB200_ORDER_DELAY). */
int hx_plan_get_processing_order(hx_plan *plan, uint32_t *order /*[C]*/);
/* ordered scatter: per cell (in processing order) the preceding cells it must wait for; pass NULLs to query nnz. */
int hx_plan_get_wait_lists(hx_plan *plan, uint32_t *nnz, uint32_t *offsets /*[C+1]*/, uint32_t *preds /*[nnz]*/);
/* Chebyshev epilogue fusion: rows whose recurrence update is applied by their last toucher inside the cell kernel
 * (owned classical, unconstrained, no hanging-node children, no halo contribution, <= 8 incident cells), and the
 * remaining owned rows that the row-list pass updates. */
int hx_plan_get_fusable_rows(hx_plan *plan, uint32_t *n_fusable, uint32_t *n_other_owned);
int hx_plan_get_c2p_transpose(hx_plan *plan, uint32_t *n_parents, uint32_t *parent_ids, uint32_t *offsets,
                              uint32_t *child_rows, double *weights); /* pass NULLs to query n_parents */

/* ---- ghost communicator (src/utils/MPICommunicatorP2P.t.cpp:77-273, 278-470) ---- */
int hx_update_ghost_values(hx_plan *plan, double *X, uint32_t B);
int hx_accumulate_add_locally_owned(hx_plan *plan, double *Y, uint32_t B);
/* ---- constraints (src/basis/ConstraintsInternal.cpp:35-108, 110-170, 172-196) ---- */
int hx_distribute_parent_to_child(hx_plan *plan, double *X, uint32_t B);
int hx_distribute_child_to_parent(hx_plan *plan, double *Y, uint32_t B);
int hx_set_constrained_nodes_to_zero(hx_plan *plan, double *Y, uint32_t B);

/* Further ConstraintsLocal objects on the same DoF numbering (set 0 = the mesh descriptor's): e.g. the inhomogeneous
 * Dirichlet constraints of the Poisson problem's X basis manager next to the homogeneous ones of Y
 * (src/electrostatics/LaplaceOperatorContextFE.t.cpp:421-433).  Same six arrays as hx_mesh_desc; returns the set id. */
int hx_plan_add_constraints(hx_plan *plan, uint32_t n_rows, const uint32_t *row_ids, const uint32_t *row_sizes,
                            const uint32_t *row_offsets, const uint32_t *col_ids, const double *col_vals,
                            const double *inhom, uint32_t *set_id);
int hx_distribute_parent_to_child_set(hx_plan *plan, uint32_t set, double *X, uint32_t B);
int hx_distribute_child_to_parent_set(hx_plan *plan, uint32_t set, double *Y, uint32_t B);

/* ---- operators: linearAlgebra::OperatorContext<double,double,DEVICE> (src/linearAlgebra/OperatorContext.h:48-111) ---- */
/* Cell-matrix operator: KohnShamOperatorContextFE (src/ksdft/KohnShamOperatorContextFE.h:102-127); also the
 * non-lumped overlap operators (src/basis/OrthoEFEOverlapOperatorContext.t.cpp:2237-2290,
 * src/basis/CFEOverlapOperatorContext.t.cpp:539-660) which run the same path with cell mass matrices. */
int hx_cellop_create(hx_plan *plan, hx_op **op);
/* reinit (src/ksdft/KohnShamOperatorContextFE.t.cpp:1240-1311): cell matrices concatenated, each n_c x n_c
 * row-major, S2 doubles; `on_device` says where `cell_matrices` lives.  Re-tiled internally. */
int hx_cellop_set_matrices(hx_op *op, const double *cell_matrices, int on_device);
int hx_cellop_set_nonlocal(hx_op *op, const hx_nonlocal_desc *nl);
/* Cells whose matrices are bitwise identical (the Laplace matrices of equally sized cells of a hexahedral mesh,
 * src/electrostatics/LaplaceOperatorContextFE.t.cpp:89-260) can share ONE re-tiled copy, which the cell kernel then
 * streams from L2 instead of HBM.  Off by default (a Kohn-Sham Hamiltonian differs cell by cell); call before
 * hx_cellop_set_matrices.  Detection = 128-bit fingerprint per cell + bitwise verification on the device; results are
 * unchanged bit for bit.  hx_cellop_num_unique_matrices reports the number of distinct matrices kept. */
int hx_cellop_set_matrix_sharing(hx_op *op, int enable);
int hx_cellop_num_unique_matrices(hx_op *op, uint32_t *n);
/* electrostatics::LaplaceOperatorContextFE (src/electrostatics/LaplaceOperatorContextFE.t.cpp:395-470): the same
 * gather -> cell GEMM -> scatter path with the grad N_i . grad N_j cell matrices, where X is filled through the
 * constraints of feBasisManagerX (x_set) and Y condensed through those of feBasisManagerY (y_set). */
int hx_cellop_set_constraint_sets(hx_op *op, uint32_t x_set, uint32_t y_set);
/* Mass-lumped M / M^-1 style operator: diagonal over local rows + atom-block enrichment matrix
 * (nE_owned x nE_owned, column-major; may be NULL when nE_owned == 0). Host pointers. */
int hx_diagop_create(hx_plan *plan, const double *diag, const double *enr_block, int variant, hx_op **op);
/* OrthoEFEOverlapInverseOpContextGLL (src/basis/OrthoEFEOverlapInverseOpContextGLL.t.cpp:1182-1282): M^-1 = diag_inv on
 * every local row + enr_block_global (nE_global x nE_global, column-major) on the enrichment rows of the WHOLE system.
 * apply(): every rank places its owned enrichment rows (global enrichment indices owned_enr_offset ..) into an
 * nE_global x B vector, the vector is all-reduced over the plan's communicator (the reference's MPI_Allreduce, :1228-1234),
 * and each rank keeps its own rows of block . Xenr; then ghost update, child->parent condensation. */
int hx_diagop_create_global_enrichment(hx_plan *plan, const double *diag_inv, const double *enr_block_global, uint32_t nE_global,
                                       uint32_t owned_enr_offset, hx_op **op);
int hx_op_destroy(hx_op *op);
/* OperatorContext::apply(X, Y, updateGhostX, updateGhostY): X may be modified (ghost update + hanging-node
 * fill, OperatorContext.h:98-101); Y fully overwritten; Y's owned rows final, ghost rows hold the rank's
 * partial sums unless updateGhostY (src/ksdft/KohnShamOperatorContextFE.t.cpp:1313-1443). */
int hx_op_apply(hx_op *op, double *X, double *Y, uint32_t B, int updateGhostX, int updateGhostY);
/* Same call with HOST buffers (n_local*B doubles each): copies X in, applies, copies X (modified) and Y out. */
int hx_op_apply_host(hx_op *op, double *X_host, double *Y_host, uint32_t B, int updateGhostX, int updateGhostY);

/* ---- preconditioned conjugate gradients: linearAlgebra::CGLinearSolver::solve (src/linearAlgebra/CGLinearSolver.t.cpp:
 * 68-300) over B right-hand sides at once with per-column step lengths, as electrostatics::PoissonLinearSolverFunctionFE
 * drives it.  b, x: DEVICE block vectors (x in: initial guess, out: xConverged); A.apply(.., true, true),
 * PC.apply(.., false, false) exactly like the reference.  status: hx_cg_status. */
int hx_cg_solve(hx_op *A, hx_op *PC, const double *b, double *x, uint32_t B, uint32_t max_iter, double abs_tol,
                double rel_tol, double div_tol, uint32_t *iterations, int *status, double *residual_norms_host);

/* ---- Chebyshev filters (src/linearAlgebra/ChebyshevFilter.h:54-113, ChebyshevFilter.t.cpp:39-134, 242-445) ---- */
/* X = eigenSubspaceGuess (in/out), Y = filteredSubspace (out); on return both hold the filtered block. */
int hx_chebyshev_filter(hx_op *A, hx_op *BInv, double *X, double *Y, uint32_t B, uint32_t degree,
                        double wantedLower, double wantedUpper, double unwantedUpper);
/* Same call with HOST buffers (n_local*B doubles each; pinned memory gives full PCIe rate): X is copied in once,
 * all `degree` applications run on the device, the filtered block is copied out into Y_host (and into X_host
 * too when write_back_x != 0, the reference's final `eigenSubspaceGuess = filteredSubspace`,
 * ChebyshevFilter.t.cpp:133).  Blocks until the result is in host memory. */
int hx_chebyshev_filter_host(hx_op *A, hx_op *BInv, double *X_host, double *Y_host, uint32_t B, uint32_t degree,
                             double wantedLower, double wantedUpper, double unwantedUpper, int write_back_x);
/* The same for a block that lives in host memory as n_batches column batches of B vectors each (Xh[k], Yh[k]: n_local x B,
 * contiguous, ideally pinned) - the column batches of ChebyshevFilteredEigenSolver::solve
 * (src/linearAlgebra/ChebyshevFilteredEigenSolver.t.cpp:231-335).  The copy-in of batch k+1 and the copy-out of batch k-1
 * overlap the filter of batch k (three streams, two device buffers per direction). */
int hx_chebyshev_filter_host_batches(hx_op *A, hx_op *BInv, const double *const *Xh, double *const *Yh, uint32_t n_batches,
                                     uint32_t B, uint32_t degree, double a0, double a, double b);
/* eigenvalues: HOST array of B doubles (std::vector<RealType>& in the reference). */
int hx_residual_chebyshev_filter(hx_op *A, hx_op *Bop, hx_op *BInv, const double *eigenvalues, double *X,
                                 double *Y, uint32_t B, uint32_t degree, double wantedLower,
                                 double wantedUpper, double unwantedUpper);

/* ---- subspace projections ---- */
/* computeXTransOpX (src/linearAlgebra/RayleighRitzEigenSolver.t.cpp:685-844 and the identical copy in
 * OrthonormalizationFunctions.t.cpp:1419-1581): column batches of `batch`, Op.apply(batch,true,false),
 * lower-trapezoid Gram block, sum over ranks.  S_host: B x B column-major, lower triangle written,
 * strict upper triangle zero. */
int hx_xtopx(hx_op *op, double *X, uint32_t B, uint32_t batch, double *S_host);
/* elpaScalaOpInternal::subspaceRotation (src/linearAlgebra/ElpaScalapackOperations.t.cpp:185-335):
 * X[dof,:] <- X[dof,:] * Q (rotationMatTranspose) or * Q^T; Q HOST, B x B column-major (replicated). */
int hx_subspace_rotation(hx_plan *plan, double *X, uint32_t B, const double *Q_host, int rotationMatTranspose,
                         int isRotationMatLowerTria);
/* MultiVector::l2Norms (src/linearAlgebra/MultiVector.t.cpp:553-578): owned rows, summed over ranks. */
int hx_l2_norms(hx_plan *plan, const double *X, uint32_t B, double *norms_host);
/* blasLapack::axpby / axpbyBlocked over the first n_rows rows (src/linearAlgebra/BlasLapackKernels.cpp:456-502). */
int hx_axpby(hx_plan *plan, uint32_t n_rows, uint32_t B, double alpha, const double *x, double beta,
             const double *y, double *z);
int hx_axpby_blocked(hx_plan *plan, uint32_t n_rows, uint32_t B, double alpha1, const double *alpha_host,
                     const double *x, double beta1, const double *beta_host, const double *y, double *z);

/* ---- the eigensolve around the path, device-resident (8f rank 4) ---- */
/* computeXTransOpX with the projected matrix left on the DEVICE: S_dev is B x B column-major, lower triangle written,
 * strict upper triangle zero, summed over ranks (no D2H / host scatter, RayleighRitzEigenSolver.t.cpp:798-836). */
int hx_xtopx_device(hx_op *op, double *X, uint32_t B, uint32_t batch, double *S_dev);
/* subspaceRotation with a DEVICE rotation matrix (column-major B x B, replicated on every rank). */
int hx_subspace_rotation_device(hx_plan *plan, double *X, uint32_t B, const double *Q_dev, int rotationMatTranspose,
                                int isRotationMatLowerTria);
/* Dense B x B steps on the device (cuSOLVER potrf + trtri / syevd on the plan's stream), the counterpart of
 * elpa_cholesky + ScaLAPACKMatrix::invert (src/linearAlgebra/OrthonormalizationFunctions.t.cpp:204-309) and
 * elpa_eigenvectors (src/linearAlgebra/RayleighRitzEigenSolver.t.cpp:154-164).
 *   cholesky_inverse: S_dev (lower triangle of an SPD matrix) -> L^-1 with S = L L^T (lower, strict upper zero);
 *   sym_eig:          S_dev (lower triangle of a symmetric matrix) -> eigenvectors in columns, eigenvalues ascending.
 * info: LAPACK-style (0 = success). */
int hx_dense_cholesky_inverse(hx_plan *plan, double *S_dev, uint32_t B, int *info);
int hx_dense_sym_eig(hx_plan *plan, double *S_dev, uint32_t B, double *eigenvalues_host, int *info);
/* OrthonormalizationFunctions::CholeskyGramSchmidt(X, orthogonalizedX, B) (OrthonormalizationFunctions.t.cpp:154-352):
 * S = X^T (Bop X), S = L L^T, X <- X L^-T in place, orthogonalizedX = X.  status: OrthonormalizationErrorCode
 * (0 SUCCESS, 1 LAPACK_ERROR, 2 NON_ORTHONORMALIZABLE_MULTIVECTOR). */
int hx_cholesky_gram_schmidt(hx_op *Bop, double *X, double *orthogonalizedX, uint32_t B, uint32_t batch, int *status);
/* RayleighRitzEigenSolver::solve(A, X, eigenValues, eigenVectors, computeEigenVectors), standard eigenproblem for an
 * M-orthonormal X (RayleighRitzEigenSolver.t.cpp:70-290): X is rotated in place, eigenVectors = X.  status 0 / 1. */
int hx_rayleigh_ritz(hx_op *A, double *X, double *eigenVectors, uint32_t B, uint32_t batch, double *eigenvalues_host,
                     int computeEigenVectors, int *status);
/* ChebyshevFilteredEigenSolver::solve (ChebyshevFilteredEigenSolver.t.cpp:189-438): filter eigenSubspaceGuess in
 * column batches of `batch` (plain or residual filter; the latter reads eigenvalues_host), Cholesky-Gram-Schmidt,
 * Rayleigh-Ritz.  On return eigenvalues_host holds the B Ritz values (ascending), eigenVectors the Ritz vectors,
 * eigenSubspaceGuess the same block (the next pass's guess).  status: EigenSolverErrorCode (0 SUCCESS,
 * 4 CHFSI_ORTHONORMALIZATION_ERROR, 5 CHFSI_RAYLEIGH_RITZ_ERROR).  B <= max_block. */
int hx_chfsi_solve(hx_op *A, hx_op *Bop, hx_op *BInv, double *eigenSubspaceGuess, double *eigenVectors, uint32_t B,
                   uint32_t batch, uint32_t degree, double wantedLower, double wantedUpper, double unwantedUpper,
                   int residualFilter, double *eigenvalues_host, int computeEigenVectors, int *status);
/* The same with the orthogonalisation of the reference's constructor argument (ChebyshevFilteredEigenSolver.t.cpp:350-371):
 * HX_ORTHO_CHOLESKY_GRAMSCHMIDT (hx_chfsi_solve) or HX_ORTHO_MULTIPASS_CGS with MultiPassOrthoDefaults. */
#define HX_ORTHO_CHOLESKY_GRAMSCHMIDT 0
#define HX_ORTHO_MULTIPASS_CGS 1
int hx_chfsi_solve_ortho(hx_op *A, hx_op *Bop, hx_op *BInv, double *eigenSubspaceGuess, double *eigenVectors, uint32_t B,
                         uint32_t batch, uint32_t degree, double wantedLower, double wantedUpper, double unwantedUpper,
                         int residualFilter, double *eigenvalues_host, int computeEigenVectors, int orthoType, int *status);
/* OrthonormalizationFunctions::MultipassCGS (src/linearAlgebra/OrthonormalizationFunctions.t.cpp:440-785): repeated
 * Cholesky-Gram-Schmidt passes with the diagonal of X^T B X shifted while its smallest eigenvalue is below shiftTolerance.
 * status: OrthonormalizationErrorCode (0 SUCCESS, 1 LAPACK_ERROR, 2 NON_ORTHONORMALIZABLE_MULTIVECTOR, 3 MAX_PASS_EXCEEDED);
 * passes (optional): Cholesky passes performed. */
int hx_multipass_cgs(hx_op *Bop, double *X, double *orthogonalizedX, uint32_t B, uint32_t batch, uint32_t maxPass,
                     double shiftTolerance, double identityTolerance, int *status, uint32_t *passes);
/* KohnShamEigenSolver::getLinearEigenSolveResidual (src/ksdft/KohnShamEigenSolver.t.cpp:574-682):
 * norms[j] = || H x_j - lambda_j M x_j ||_2 over owned rows, in column batches. */
int hx_eigen_residual_norms(hx_op *A, hx_op *Mop, const double *X, uint32_t B, uint32_t batch,
                            const double *eigenvalues_host, double *norms_host);
/* LanczosExtremeEigenSolver::solve, eigenvalues only (src/linearAlgebra/LanczosExtremeEigenSolver.t.cpp:216-520):
 * B-orthogonal Lanczos on BInv A from initialGuess (DEVICE, n_local doubles, one vector).  eigenvalues_host:
 * numLower lowest then numUpper highest Ritz values; diagonal/subDiagonal_host (maxKrylovSubspaceSize doubles each,
 * may be NULL) receive the tridiagonal matrix (getTridiagonalMatrix), krylovSize its order.  status:
 * EigenSolverErrorCode (0 SUCCESS, 1 LAPACK_ERROR, 2 LANCZOS_BETA_ZERO, 3 LANCZOS_SUBSPACE_INSUFFICIENT, 11 OTHER). */
int hx_lanczos_extreme(hx_op *A, hx_op *Bop, hx_op *BInv, const double *initialGuess, uint32_t maxKrylovSubspaceSize,
                       uint32_t numLower, uint32_t numUpper, const double *tolerance, double lanczosBetaTolerance,
                       int adaptive, double *eigenvalues_host, double *diagonal_host, double *subDiagonal_host,
                       uint32_t *krylovSize, int *status);

/* ---- cell-matrix assembly, the step before the path each SCF iteration ---- */
/* The per-cell data FEBasisOperations::computeFEMatrices reads from FEBasisDataStorage
 * (src/basis/FEBasisOperations.t.cpp:111-160): quadrature points per cell, JxW, and the basis values per cell as
 * getBasisDataInCellRange lays them out (nq_c x n_c, DoF index fastest).  HOST pointers, copied at creation.
 * same_basis_in_all_cells = the reference's zeroStrideBasisVal (same quadrature rule in every cell and a fixed number of
 * DoFs per cell, :133-137): basis_data then holds ONE matrix. */
typedef struct hx_fe_basis hx_fe_basis;
typedef struct hx_fe_basis_desc {
  uint32_t        struct_size;
  int32_t         same_basis_in_all_cells;
  const uint32_t *num_cell_quad; /* nCellQuadraturePoints(c)  [C]                       */
  const double *  basis_data;    /* concatenated per cell, or one matrix                */
  const double *  jxw;           /* getJxWInAllCells()        [sum nq_c]                */
} hx_fe_basis_desc;
int hx_fe_basis_create(hx_plan *plan, const hx_fe_basis_desc *desc, hx_fe_basis **basis);
int hx_fe_basis_destroy(hx_fe_basis *basis);
/* FEBasisOperations::computeFEMatrices(IDENTITY, MULT, MULT, IDENTITY, f, cellWiseFEData)
 * (src/basis/FEBasisOperations.t.cpp:2210-2243, 41-427): cell_matrices_dev[c] (n_c x n_c) =
 * sum_q N[q,i] f[q] JxW[q] N[q,j] (+ add_to_dev[c] when given: the component sum of KohnShamOperatorContextFE::reinit,
 * src/ksdft/KohnShamOperatorContextFE.t.cpp:1259-1282).  f: one value per quadrature point (host or device);
 * output: DEVICE, S2 doubles, the layout hx_cellop_set_matrices(.., on_device = 1) takes. */
int hx_compute_fe_matrices(hx_fe_basis *basis, const double *f_quad, int f_on_device, const double *add_to_dev,
                           double *cell_matrices_dev);

/* The same computation written straight into the operator's re-tiled matrix stream: computeFEMatrices + the component
 * sum + KohnShamOperatorContextFE::reinit in ONE kernel (no flat S2 array, no re-tiling pass).  The operator keeps its
 * nonlocal part (hx_cellop_set_nonlocal); add_to_dev (flat S2 layout, e.g. the kinetic cell matrices) may be NULL. */
int hx_cellop_assemble_matrices(hx_op *op, hx_fe_basis *basis, const double *f_quad, int f_on_device,
                                const double *add_to_dev);

/* ---- density, the step after the path each SCF iteration ---- */
/* DensityCalculator::computeRho(occupation, waveFunc, rho) (src/ksdft/DensityCalculator.t.cpp:283-437):
 * rho[q] = sum_i 2 occupation[i] |psi_i(q)|^2 with psi_i(q) = sum_j N_c[q,j] X[cellLocalIds_c[j], i]
 * (FEBasisOperations::interpolate, src/basis/FEBasisOperations.t.cpp:996-1275, then computeRhoInBatch, :37-70).
 * X is used as it stands - like the reference, no ghost update or hanging-node fill happens here (the eigensolver
 * leaves the wavefunctions ghost-updated, src/ksdft/KohnShamEigenSolver.t.cpp:333).  rho: one value per quadrature
 * point, host or device. */
int hx_compute_rho(hx_fe_basis *basis, const double *X_dev, uint32_t B, const double *occupation_host, double *rho,
                   int rho_on_device);

/* Chebyshev polynomial degree for a spectral upper bound: LinearEigenSolverDefaults::CHEBY_ORDER_LOOKUP through
 * getChebyPolynomialDegree (src/ksdft/Defaults.cpp:51-58, src/ksdft/KohnShamEigenSolver.t.cpp:37-46). */
int hx_chebyshev_polynomial_degree(double unWantedSpectrumUpperBound, uint32_t *degree);

/* ---- measurement hooks (bench.py / tests) ---- */
/* Number of kernel launches issued through this plan since creation, and device time of the dominant
 * cell-contraction kernel accumulated with CUDA events on the plan's stream (reset on read). */
int hx_plan_launch_count(hx_plan *plan, uint64_t *n);
int hx_plan_cell_kernel_time_ms(hx_plan *plan, double *ms, uint64_t *launches);
int hx_plan_enable_kernel_timing(hx_plan *plan, int on);
/* SM clock the cell kernel actually ran at while kernel timing was on: clock64 cycles / globaltimer nanoseconds of its
 * CTA 0, summed over the launches (reset on read).  Shows power-cap throttling that a 1 Hz nvidia-smi sample misses. */
int hx_plan_cell_kernel_sm_clock_mhz(hx_plan *plan, double *mhz);
/* Phase trace: CUDA events at the phase boundaries of every apply / filter degree issued while tracing is on;
 * the report is a JSON object {"phase": {"ms": total, "n": count}, ...} (phases: x-halo, p2c+zero, nl-phase-a,
 * nl-halo, cell-kernel, shared+c2p, y-halo, cheb-rest) and clears the trace. */
int hx_plan_trace(hx_plan *plan, int on);
int hx_plan_trace_report(hx_plan *plan, char *buf, size_t buf_bytes);
/* 1 when the short kernels of an apply / filter degree are launched with programmatic dependent launch (the launch of
 * kernel k+1 overlaps the tail of kernel k; results are bitwise those of serialised launches).  Default: on
 * (single- and multi-rank plans); environment HXB200_PDL=0 switches it off
 * (read at every launch).  No reference counterpart (the reference launches nothing). */
int hx_programmatic_launch_enabled(void);
/* FP64 DMMA / DFMA / copy microbenchmarks used for the roofline denominators. */
int hx_microbench(double *dmma_tflops, double *dfma_tflops, double *copy_gbs);

#ifdef __cplusplus
}
#endif
#endif /* HXB200_H */
