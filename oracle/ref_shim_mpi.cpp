// ref_shim_mpi.cpp — serial definitions of the five utils::mpi wrappers the reference's leaf
// sources reference.  TEST INFRASTRUCTURE ONLY.  The reference's own serial branch
// (src/utils/MPIWrapper.cpp:188-361) does not compile (stray ';' at MPIWrapper.cpp:334), and MPI
// is absent here, so oracle/_ref is a single-rank build: communicator size 1, rank 0, every
// request complete.
#include <string>
#include <utility>
#include <utils/MPITypes.h>
#include <utils/MPIWrapper.h>
namespace dftefe
{
  namespace utils
  {
    namespace mpi
    {
      int
      MPICommRank(MPIComm, int *rank)
      {
        *rank = 0;
        return MPISuccess;
      }
      int
      MPICommSize(MPIComm, int *size)
      {
        *size = 1;
        return MPISuccess;
      }
      int
      MPIBarrier(MPIComm)
      {
        return MPISuccess;
      }
      std::pair<bool, std::string>
      MPIErrIsSuccessAndMsg(int errCode)
      {
        return std::make_pair(errCode == MPISuccess, std::string(errCode == MPISuccess ? "" : "serial MPI stub error"));
      }
      int
      MPIIbarrier(MPIComm, MPIRequest *)
      {
        return MPISuccess;
      }
      int
      MPITest(MPIRequest *, int *flag, MPIStatus *)
      {
        *flag = 1;
        return MPISuccess;
      }
      int
      MPITestall(int, MPIRequest *, int *flag, MPIStatus *)
      {
        *flag = 1;
        return MPISuccess;
      }
    } // namespace mpi
  }   // namespace utils
} // namespace dftefe
