// api.cu — host-side logic of libhxb200 and the extern "C" entry points declared in include/hxb200.h.
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <map>
#include <memory>
#include <queue>
#include <new>

#include "hx_internal.h"

namespace hx
{
  static thread_local char g_err[1024] = "";
  void
  set_error(const char *fmt, ...)
  {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
  }

  // Programmatic dependent launch of the kernels of an apply (read per launch so a test can toggle it): on by default,
  // HXB200_PDL=0 switches it off.  (Multi-rank plans with the spin-waiting peer-memory halo kernels were validated with it
  // on 2 B200s in round 2: tests/test_multi_gpu.py under HXB200_PDL=1, profiles/r2_mgpu_parity_N2_serial_halo.log.)
  bool
  pdl_enabled()
  {
    const char *e = getenv("HXB200_PDL");
    return !(e && e[0] == '0');
  }

  int
  Halo::init(const hx_halo_desc &h, uint32_t max_block)
  {
    n_owned = h.n_owned;
    n_ghost = h.n_ghost;
    ghost_procs.assign(h.ghost_proc_ids, h.ghost_proc_ids + h.n_ghost_procs);
    ghost_ranges.assign(h.ghost_ranges, h.ghost_ranges + 2 * (size_t)h.n_ghost_procs);
    target_procs.assign(h.target_proc_ids, h.target_proc_ids + h.n_target_procs);
    target_counts.assign(h.num_owned_for_target, h.num_owned_for_target + h.n_target_procs);
    n_send = 0;
    for (uint32_t c : target_counts)
      n_send += c;
    uint32_t nrecv = 0;
    for (uint32_t i = 0; i < h.n_ghost_procs; ++i)
      {
        HX_CHECK(ghost_ranges[2 * i + 1] >= ghost_ranges[2 * i], HX_ERR_INVALID, "halo: bad ghost range %u", i);
        nrecv += ghost_ranges[2 * i + 1] - ghost_ranges[2 * i];
      }
    HX_CHECK(nrecv == n_ghost, HX_ERR_INVALID, "halo: ghost ranges cover %u of %u ghosts", nrecv, n_ghost);
    for (uint32_t i = 0; i < n_ghost; ++i)
      HX_CHECK(h.ghost_local_ids[i] < n_ghost, HX_ERR_INVALID, "halo: ghost local id out of range");
    for (uint32_t i = 0; i < n_send; ++i)
      HX_CHECK(h.owned_local_ids_for_targets[i] < n_owned, HX_ERR_INVALID, "halo: owned id for target out of range");
    HX_TRY(d_ghost_local_ids.upload(h.ghost_local_ids, n_ghost));
    HX_TRY(d_owned_ids_for_targets.upload(h.owned_local_ids_for_targets, n_send));
    // accumulate CSR: owned row -> positions in the receive buffer, ascending (= reference buffer order)
    std::vector<std::pair<uint32_t, uint32_t>> rp(n_send);
    for (uint32_t i = 0; i < n_send; ++i)
      rp[i] = {h.owned_local_ids_for_targets[i], i};
    std::stable_sort(rp.begin(), rp.end(), [](auto &x, auto &y) { return x.first < y.first; });
    std::vector<uint32_t> rows, off, pos(n_send);
    for (uint32_t i = 0; i < n_send; ++i)
      {
        if (i == 0 || rp[i].first != rp[i - 1].first)
          {
            rows.push_back(rp[i].first);
            off.push_back(i);
          }
        pos[i] = rp[i].second;
      }
    off.push_back(n_send);
    n_acc_rows = (uint32_t)rows.size();
    HX_TRY(d_acc_rows.upload(rows));
    HX_TRY(d_acc_off.upload(off));
    HX_TRY(d_acc_pos.upload(pos));
    buf_doubles = (size_t)std::max(n_send, n_ghost) * max_block; // staging buffers of the NCCL transport: allocated on its first use
    return HX_OK;
  }

  // first use of a halo after the communicator is attached: bring up the NVLink peer-memory transport (collective;
  // every rank reaches this point for the same halo), or stay on NCCL send/recv (HXB200_HALO=nccl, or no IPC)
  static int
  halo_pick_transport(hx_plan *p, Halo &h)
  {
    if (h.peer || h.peer_failed)
      return HX_OK;
    const char *e = getenv("HXB200_HALO");
    if (e && strcmp(e, "nccl") == 0)
      h.peer_failed = true;
    else
      HX_TRY(peer_setup(p, h));
    if (&h == &p->halo)
      p->halo_transport = h.peer ? 2 : 1;
    return HX_OK;
  }

  // collective teardown of a peer transport: no rank may unmap / free its arena while a neighbour's last
  // acknowledgement can still be in flight
  static void
  halo_peer_release(hx_plan *p, Halo &h)
  {
    if (!h.peer)
      return;
    cudaStreamSynchronize(p->stream);
    if (p->comm)
      {
        std::vector<unsigned char> f1(4, 0), fall(4 * (size_t)p->nranks, 0);
        comm_allgather_bytes(p->comm, p->stream, f1.data(), fall.data(), 4);
      }
    p->peer_halos.erase(std::remove(p->peer_halos.begin(), p->peer_halos.end(), &h), p->peer_halos.end());
    peer_destroy(h.peer);
    h.peer = nullptr;
  }

  // every synchronisation that hands results to the host goes through here: a peer halo wait that timed out inside a
  // kernel (lost neighbour) must surface as HX_ERR_COMM, not as host scalars computed from garbage
  int
  plan_sync(hx_plan *p)
  {
    HX_CUDA(cudaStreamSynchronize(p->stream));
    for (Halo *h : p->peer_halos)
      HX_TRY(peer_check_status(*h));
    return HX_OK;
  }

  int
  halo_update(hx_plan *p, Halo &h, double *X, uint32_t B)
  {
    if (p->nranks == 1)
      return HX_OK;
    HX_CHECK(p->comm != nullptr, HX_ERR_COMM, "nranks > 1 but no communicator attached (hx_plan_attach_comm)");
    HX_TRY(halo_pick_transport(p, h)); // collective: before the early-out of ranks with nothing to exchange
    if (h.n_ghost == 0 && h.n_send == 0)
      return HX_OK;
    if (h.peer)
      return peer_halo_update(p, h, X, B);
    HX_TRY(h.ensure_staging());
    HX_TRY(launch_pack(p, X, B, h.d_owned_ids_for_targets.p, h.n_send, h.d_send.p));
    std::vector<size_t> sc(h.target_counts.size()), rc(h.ghost_procs.size());
    for (size_t i = 0; i < sc.size(); ++i)
      sc[i] = (size_t)h.target_counts[i] * B;
    for (size_t i = 0; i < rc.size(); ++i)
      rc[i] = (size_t)(h.ghost_ranges[2 * i + 1] - h.ghost_ranges[2 * i]) * B;
    HX_TRY(comm_exchange(p->comm, p->stream, h.d_send.p, h.target_procs, sc, h.d_recv.p, h.ghost_procs, rc));
    return launch_unpack(p, h.d_recv.p, B, h.d_ghost_local_ids.p, h.n_ghost, X + (size_t)h.n_owned * B);
  }

  static int
  halo_accumulate(hx_plan *p, Halo &h, double *Y, uint32_t B)
  {
    if (p->nranks == 1)
      return HX_OK;
    HX_CHECK(p->comm != nullptr, HX_ERR_COMM, "nranks > 1 but no communicator attached (hx_plan_attach_comm)");
    HX_TRY(halo_pick_transport(p, h));
    if (h.n_ghost == 0 && h.n_send == 0)
      return HX_OK;
    if (h.peer)
      return peer_halo_accumulate(p, h, Y, B);
    HX_TRY(h.ensure_staging());
    HX_TRY(launch_pack(p, Y + (size_t)h.n_owned * B, B, h.d_ghost_local_ids.p, h.n_ghost, h.d_send.p));
    std::vector<size_t> sc(h.ghost_procs.size()), rc(h.target_counts.size());
    for (size_t i = 0; i < sc.size(); ++i)
      sc[i] = (size_t)(h.ghost_ranges[2 * i + 1] - h.ghost_ranges[2 * i]) * B;
    for (size_t i = 0; i < rc.size(); ++i)
      rc[i] = (size_t)h.target_counts[i] * B;
    HX_TRY(comm_exchange(p->comm, p->stream, h.d_send.p, h.ghost_procs, sc, h.d_recv.p, h.target_procs, rc));
    return launch_add_rows(p, h.d_recv.p, B, h.d_acc_rows.p, h.d_acc_off.p, h.d_acc_pos.p, h.n_acc_rows, Y);
  }

  // accumulateAddLocallyOwned + updateGhostValues of the same vector (a sum over the sharing ranks that everybody needs):
  // one single-block kernel when the halo is small and the peer transport is up (HXB200_NL_FUSED=0: the two exchanges)
  static int
  halo_accumulate_update(hx_plan *p, Halo &h, double *X, uint32_t B)
  {
    if (p->nranks == 1)
      return HX_OK;
    HX_CHECK(p->comm != nullptr, HX_ERR_COMM, "nranks > 1 but no communicator attached (hx_plan_attach_comm)");
    HX_TRY(halo_pick_transport(p, h));
    if (h.n_ghost == 0 && h.n_send == 0)
      return HX_OK;
    static const bool fused_ok = !(getenv("HXB200_NL_FUSED") && getenv("HXB200_NL_FUSED")[0] == '0');
    if (fused_ok && peer_acc_update_is_small(h, B))
      return peer_halo_accumulate_update_small(p, h, X, B);
    HX_TRY(halo_accumulate(p, h, X, B));
    return halo_update(p, h, X, B);
  }

  // ---------------------------------------------------------------------------------------------
  static int
  sm_count_for_order()
  {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess)
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }

  static int
  build_plan(hx_plan *p, const hx_mesh_desc *m)
  {
    p->rank              = m->rank;
    p->nranks            = m->nranks;
    p->n_owned           = m->halo.n_owned;
    p->n_ghost           = m->halo.n_ghost;
    p->n_local           = p->n_owned + p->n_ghost;
    p->n_owned_classical = m->n_owned_classical;
    p->C                 = m->n_cells;
    p->max_block         = m->max_block ? m->max_block : 1;
    HX_CHECK(p->n_owned_classical <= p->n_owned, HX_ERR_INVALID, "n_owned_classical > n_owned");
    HX_CHECK(p->nranks >= 1 && p->rank >= 0 && p->rank < p->nranks, HX_ERR_INVALID, "bad rank/nranks");
    HX_CHECK(p->n_local < 0x0fffffffu, HX_ERR_INVALID, "too many local rows");

    p->h_ncd.assign(m->num_cell_dofs, m->num_cell_dofs + p->C);
    p->h_cell_off.assign(p->C + 1, 0);
    p->S2    = 0;
    p->max_n = 0;
    for (uint32_t c = 0; c < p->C; ++c)
      {
        HX_CHECK((uint64_t)p->h_cell_off[c] + p->h_ncd[c] < 0xffffffffull, HX_ERR_INVALID, "cell map too large");
        p->h_cell_off[c + 1] = p->h_cell_off[c] + p->h_ncd[c];
        p->S2 += (size_t)p->h_ncd[c] * p->h_ncd[c];
        p->max_n = std::max(p->max_n, p->h_ncd[c]);
      }
    p->S = p->h_cell_off[p->C];
    p->h_ids.assign(m->cell_local_ids, m->cell_local_ids + p->S);
    for (uint32_t i = 0; i < p->S; ++i)
      HX_CHECK(p->h_ids[i] < p->n_local, HX_ERR_INVALID, "cell_local_ids[%u] = %u out of range", i, p->h_ids[i]);
    HX_TRY(p->d_ids.upload(p->h_ids));
    HX_TRY(p->d_cell_off.upload(p->h_cell_off));
    HX_TRY(p->d_ncd.upload(p->h_ncd));

    // ---- constraints ----
    p->nR  = m->n_constraint_rows;
    p->nnz = 0;
    std::vector<uint32_t> rowinfo(p->n_local, 0xFFFFFFFFu);
    for (uint32_t i = 0; i < p->nR; ++i)
      {
        HX_CHECK(m->row_ids[i] < p->n_local, HX_ERR_INVALID, "constraint row id out of range");
        HX_CHECK(rowinfo[m->row_ids[i]] == 0xFFFFFFFFu, HX_ERR_INVALID, "duplicate constraint row %u", m->row_ids[i]);
        rowinfo[m->row_ids[i]] = 0xFFFFFFFEu;
        p->nnz                 = std::max(p->nnz, m->row_offsets[i] + m->row_sizes[i]);
        p->max_row             = std::max(p->max_row, m->row_sizes[i]);
      }
    std::map<uint32_t, std::vector<std::pair<uint32_t, double>>> par;
    for (uint32_t i = 0; i < p->nR; ++i)
      for (uint32_t j = 0; j < m->row_sizes[i]; ++j)
        {
          const uint32_t col = m->col_ids[m->row_offsets[i] + j];
          HX_CHECK(col < p->n_local, HX_ERR_INVALID, "constraint column id out of range");
          // the parallel distribute kernels need closed constraints (deal.II AffineConstraints::close(),
          // reference src/basis/CFEConstraintsLocalDealii.t.cpp:96-104)
          HX_CHECK(rowinfo[col] != 0xFFFFFFFEu, HX_ERR_UNSUPPORTED,
                   "constraint chain: row %u depends on constrained row %u (constraints must be closed)",
                   m->row_ids[i], col);
          par[col].push_back({m->row_ids[i], m->col_vals[m->row_offsets[i] + j]});
        }
    p->h_par_ids.clear();
    p->h_par_off.assign(1, 0);
    p->h_par_child.clear();
    p->h_par_w.clear();
    for (auto &kv : par)
      {
        rowinfo[kv.first] = (uint32_t)p->h_par_ids.size();
        p->h_par_ids.push_back(kv.first);
        for (auto &e : kv.second)
          {
            p->h_par_child.push_back(e.first);
            p->h_par_w.push_back(e.second);
          }
        p->h_par_off.push_back((uint32_t)p->h_par_child.size());
        p->max_child = std::max(p->max_child, (uint32_t)kv.second.size());
      }
    p->nPar = (uint32_t)p->h_par_ids.size();
    // M^-1 of the fused Chebyshev step reads its input at the row itself, at the constrained children of a parent row and
    // at the owned enrichment rows: without parents and without constrained enrichment rows no constrained row is read
    p->cheb_fill_dead = (p->nPar == 0);
    for (uint32_t i = 0; i < p->nR; ++i)
      if (m->row_ids[i] >= p->n_owned_classical && m->row_ids[i] < p->n_owned)
        p->cheb_fill_dead = false;
    p->h_row_ids.assign(m->row_ids, m->row_ids + p->nR);
    HX_TRY(p->d_row_ids.upload(m->row_ids, p->nR));
    HX_TRY(p->d_row_sizes.upload(m->row_sizes, p->nR));
    HX_TRY(p->d_row_offsets.upload(m->row_offsets, p->nR));
    HX_TRY(p->d_col_ids.upload(m->col_ids, p->nnz));
    HX_TRY(p->d_col_vals.upload(m->col_vals, p->nnz));
    HX_TRY(p->d_inhom.upload(m->inhom, p->nR));
    HX_TRY(p->d_par_ids.upload(p->h_par_ids));
    HX_TRY(p->d_par_off.upload(p->h_par_off));
    HX_TRY(p->d_par_child.upload(p->h_par_child));
    HX_TRY(p->d_par_w.upload(p->h_par_w));
    HX_TRY(p->d_rowinfo.upload(rowinfo));

    // ---- shared rows: touched by more than 8 cells (enrichment DoFs); they bypass the colouring ----
    std::vector<uint32_t> incidence(p->n_local, 0);
    for (uint32_t i = 0; i < p->S; ++i)
      incidence[p->h_ids[i]]++;
    const uint32_t        SHARED_THRESHOLD = 8;
    std::vector<uint32_t> dest(p->S);
    std::map<uint32_t, std::vector<uint32_t>> shared;
    p->n_slots = 0;
    for (uint32_t i = 0; i < p->S; ++i)
      {
        const uint32_t r = p->h_ids[i];
        if (incidence[r] > SHARED_THRESHOLD)
          {
            shared[r].push_back(p->n_slots);
            dest[i] = HX_DEST_STAGED | p->n_slots;
            p->n_slots++;
          }
        else
          dest[i] = r;
      }
    std::vector<uint32_t> sh_rows, sh_off(1, 0), sh_slots;
    for (auto &kv : shared)
      {
        sh_rows.push_back(kv.first);
        sh_slots.insert(sh_slots.end(), kv.second.begin(), kv.second.end());
        sh_off.push_back((uint32_t)sh_slots.size());
      }
    p->n_shared = (uint32_t)sh_rows.size();
    HX_TRY(p->d_sh_rows.upload(sh_rows));
    HX_TRY(p->d_sh_off.upload(sh_off));
    HX_TRY(p->d_sh_slots.upload(sh_slots));
    HX_TRY(p->d_stage.alloc((size_t)p->n_slots * p->max_block));
    {
      const uint32_t SH_CHUNK = 64;
      uint32_t       max_cnt  = 0;
      for (uint32_t r = 0; r < p->n_shared; ++r)
        max_cnt = std::max(max_cnt, sh_off[r + 1] - sh_off[r]);
      p->n_sh_chunks = 0;
      if (max_cnt > 2 * SH_CHUNK)
        {
          std::vector<uint32_t> cb, ce, coff(1, 0);
          for (uint32_t r = 0; r < p->n_shared; ++r)
            {
              for (uint32_t e = sh_off[r]; e < sh_off[r + 1]; e += SH_CHUNK)
                {
                  cb.push_back(e);
                  ce.push_back(std::min(e + SH_CHUNK, sh_off[r + 1]));
                }
              coff.push_back((uint32_t)cb.size());
            }
          p->n_sh_chunks = (uint32_t)cb.size();
          HX_TRY(p->d_sh_ch_begin.upload(cb));
          HX_TRY(p->d_sh_ch_end.upload(ce));
          HX_TRY(p->d_sh_ch_off.upload(coff));
          HX_TRY(p->d_sh_partial.alloc((size_t)p->n_sh_chunks * p->max_block));
        }
    }

    // ---- greedy cell colouring on the cell-DoF graph (non-shared rows only) ----
    // cell c takes the lowest colour not used by any earlier cell sharing a row with it.
    std::vector<uint64_t> used(p->n_local, 0);
    p->h_colour.assign(p->C, 0);
    p->n_colours = 0;
    for (uint32_t c = 0; c < p->C; ++c)
      {
        uint64_t mask = 0;
        for (uint32_t i = p->h_cell_off[c]; i < p->h_cell_off[c + 1]; ++i)
          if (!(dest[i] & HX_DEST_STAGED))
            mask |= used[p->h_ids[i]];
        HX_CHECK(mask != ~0ull, HX_ERR_UNSUPPORTED, "cell %u needs more than 64 colours", c);
        uint32_t col = 0;
        while (mask & (1ull << col))
          ++col;
        p->h_colour[c] = col;
        p->n_colours   = std::max(p->n_colours, col + 1);
        for (uint32_t i = p->h_cell_off[c]; i < p->h_cell_off[c + 1]; ++i)
          if (!(dest[i] & HX_DEST_STAGED))
            used[p->h_ids[i]] |= (1ull << col);
      }
    // a row listed twice inside one cell would race inside the CTA: reject
    {
      std::vector<uint32_t> seen(p->n_local, 0xffffffffu);
      for (uint32_t c = 0; c < p->C; ++c)
        for (uint32_t i = p->h_cell_off[c]; i < p->h_cell_off[c + 1]; ++i)
          {
            HX_CHECK(seen[p->h_ids[i]] != c, HX_ERR_UNSUPPORTED, "cell %u lists local row %u twice", c, p->h_ids[i]);
            seen[p->h_ids[i]] = c;
          }
    }
    p->h_colour_off.assign(p->n_colours + 1, 0);
    for (uint32_t c = 0; c < p->C; ++c)
      p->h_colour_off[p->h_colour[c] + 1]++;
    for (uint32_t k = 0; k < p->n_colours; ++k)
      p->h_colour_off[k + 1] += p->h_colour_off[k];
    p->h_colour_cells.resize(p->C);
    {
      std::vector<uint32_t> fill(p->h_colour_off.begin(), p->h_colour_off.end() - 1);
      for (uint32_t c = 0; c < p->C; ++c)
        p->h_colour_cells[fill[p->h_colour[c]]++] = c;
    }
    HX_TRY(p->d_colour_cells.upload(p->h_colour_cells));

    // cells that read ghost rows of X (multi-rank plans)
    std::vector<char> boundary(p->C, 0);
    p->n_boundary_cells = 0;
    if (p->nranks > 1)
      for (uint32_t c = 0; c < p->C; ++c)
        {
          for (uint32_t i = p->h_cell_off[c]; i < p->h_cell_off[c + 1]; ++i)
            if (p->h_ids[i] >= p->n_owned)
              boundary[c] = 1;
          p->n_boundary_cells += boundary[c];
        }
    p->h_boundary.assign(boundary.begin(), boundary.end());

    // ---- ordered scatter: processing order, first-touch flags, predecessor (wait) lists ----
    // For every non-shared row the touching cells form a chain in processing order; a cell waits only for the
    // immediately preceding toucher of each of its rows (completion is transitive).  The order keeps the
    // touches of a row close in time (its X and Y lines stay in L2) yet at least D positions apart, so a cell
    // practically never stalls on a predecessor that is still contracting.
    {
      p->h_order.resize(p->C);
      for (uint32_t c = 0; c < p->C; ++c)
        p->h_order[c] = c;
      if (const char *e = getenv("HXB200_ORDER_BLOCK"))
        {
          // alternative: blocks of consecutive cells, each stably sorted by colour
          const uint32_t order_block = (uint32_t)std::max(1, atoi(e));
          for (uint32_t b0 = 0; b0 < p->C; b0 += order_block)
            std::stable_sort(p->h_order.begin() + b0, p->h_order.begin() + std::min(p->C, b0 + order_block),
                             [&](uint32_t x, uint32_t y) { return p->h_colour[x] < p->h_colour[y]; });
        }
      else
        {
          // delay-D list schedule: sweep the cells in the caller's order, but place a cell only when every
          // neighbour (cell sharing a non-staged row) already placed sits >= D positions back; skipped cells
          // are picked up as soon as they become eligible.  D ~ twice the number of resident CTAs.
          uint32_t D = 4u * (uint32_t)std::max(1, sm_count_for_order());
          if (const char *e2 = getenv("HXB200_ORDER_DELAY"))
            D = (uint32_t)std::max(0, atoi(e2));
          // row -> cells CSR over non-staged rows
          std::vector<uint32_t> rc_off(p->n_local + 1, 0), rc;
          for (uint32_t i = 0; i < p->S; ++i)
            if (!(dest[i] & HX_DEST_STAGED))
              rc_off[p->h_ids[i] + 1]++;
          for (uint32_t r = 0; r < p->n_local; ++r)
            rc_off[r + 1] += rc_off[r];
          rc.resize(rc_off[p->n_local]);
          {
            std::vector<uint32_t> fill(rc_off.begin(), rc_off.end() - 1);
            for (uint32_t c = 0; c < p->C; ++c)
              for (uint32_t i = p->h_cell_off[c]; i < p->h_cell_off[c + 1]; ++i)
                if (!(dest[i] & HX_DEST_STAGED))
                  rc[fill[p->h_ids[i]]++] = c;
          }
          // sweep key of a cell: its index in the caller's order - except that on a multi-rank plan the cells reading ghost
          // rows are all keyed at a quarter of the sweep: late enough that the halo has arrived when the first of them is
          // claimed, early enough that their ghost-row sums travel while the rest of the interior is contracted
          typedef unsigned long long KEY;
          const uint32_t k0 = p->C / 4;
          auto           key_of = [&](uint32_t c) -> KEY { return ((KEY)(boundary[c] ? k0 : c) << 33) | ((KEY)(boundary[c] ? 0u : 1u) << 32) | c; };
          typedef std::pair<uint32_t, KEY> PR; // (time, key)
          std::priority_queue<KEY, std::vector<KEY>, std::greater<KEY>> eligible;
          std::priority_queue<PR, std::vector<PR>, std::greater<PR>>    waiting;
          std::vector<uint32_t> ready(p->C, 0);
          std::vector<char>     placed(p->C, 0);
          for (uint32_t c = 0; c < p->C; ++c)
            eligible.push(key_of(c));
          for (uint32_t t = 0; t < p->C; ++t)
            {
              while (!waiting.empty() && waiting.top().first <= t)
                {
                  const PR       e = waiting.top();
                  const uint32_t x = (uint32_t)e.second;
                  waiting.pop();
                  if (!placed[x])
                    {
                      if (ready[x] <= t)
                        eligible.push(e.second);
                      else if (ready[x] != e.first)
                        waiting.push(PR(ready[x], e.second));
                    }
                }
              uint32_t c = 0xffffffffu;
              while (!eligible.empty())
                {
                  const KEY      k = eligible.top();
                  const uint32_t x = (uint32_t)k;
                  eligible.pop();
                  if (placed[x])
                    continue;
                  if (ready[x] > t)
                    {
                      waiting.push(PR(ready[x], k));
                      continue;
                    }
                  c = x;
                  break;
                }
              while (c == 0xffffffffu)
                {
                  // every remaining cell is delayed: take the one that becomes eligible first
                  const PR       e = waiting.top();
                  const uint32_t x = (uint32_t)e.second;
                  waiting.pop();
                  if (placed[x])
                    continue;
                  if (ready[x] != e.first)
                    {
                      waiting.push(PR(ready[x], e.second));
                      continue;
                    }
                  c = x;
                }
              placed[c]     = 1;
              p->h_order[t] = c;
              for (uint32_t i = p->h_cell_off[c]; i < p->h_cell_off[c + 1]; ++i)
                if (!(dest[i] & HX_DEST_STAGED))
                  {
                    const uint32_t r = p->h_ids[i];
                    for (uint32_t e = rc_off[r]; e < rc_off[r + 1]; ++e)
                      if (!placed[rc[e]])
                        ready[rc[e]] = t + D;
                  }
            }
        }
      std::vector<uint32_t> last(p->n_local, 0xffffffffu); // last processing index that touched the row
      std::vector<uint32_t> last_entry(p->n_local, 0xffffffffu); // its position in the cell->row map
      p->h_wait_off.assign(p->C + 1, 0);
      p->h_wait_list.clear();
      std::vector<uint32_t> tmp;
      for (uint32_t w = 0; w < p->C; ++w)
        {
          const uint32_t c = p->h_order[w];
          tmp.clear();
          for (uint32_t i = p->h_cell_off[c]; i < p->h_cell_off[c + 1]; ++i)
            {
              if (dest[i] & HX_DEST_STAGED)
                continue;
              const uint32_t r = p->h_ids[i];
              if (last[r] == 0xffffffffu)
                dest[i] |= HX_DEST_FIRST;
              else if (last[r] != w)
                tmp.push_back(last[r]);
              last[r]       = w;
              last_entry[r] = i;
            }
          std::sort(tmp.begin(), tmp.end());
          tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
          p->h_wait_list.insert(p->h_wait_list.end(), tmp.begin(), tmp.end());
          p->h_wait_off[w + 1] = (uint32_t)p->h_wait_list.size();
        }
      // Chebyshev epilogue fusion: the last toucher of a fusable row applies the recurrence (hx_internal.h)
      {
        std::vector<char> acc_target(p->n_local, 0);
        uint32_t          nsend = 0;
        for (uint32_t t = 0; t < m->halo.n_target_procs; ++t)
          nsend += m->halo.num_owned_for_target[t];
        if (p->nranks > 1)
          for (uint32_t i = 0; i < nsend; ++i)
            if (m->halo.owned_local_ids_for_targets[i] < p->n_local)
              acc_target[m->halo.owned_local_ids_for_targets[i]] = 1;
        std::vector<uint32_t> nonfuse;
        p->n_fusable = 0;
        for (uint32_t r = 0; r < p->n_owned; ++r)
          {
            const bool ok = r < p->n_owned_classical && rowinfo[r] == 0xFFFFFFFFu && !acc_target[r] &&
                            last_entry[r] != 0xffffffffu;
            if (ok)
              {
                dest[last_entry[r]] |= HX_DEST_LASTF;
                p->n_fusable++;
              }
            else
              nonfuse.push_back(r);
          }
        p->n_nonfuse = (uint32_t)nonfuse.size();
        HX_TRY(p->d_nonfuse_rows.upload(nonfuse));
        // the fused M^-1 skips the ghost update between its row scaling and its child->parent pass: exact unless a
        // constrained GHOST row has parents (its scaled value would come from the owner)
        p->cheb_fusable_multirank = true;
        for (uint32_t i = 0; i < p->nR; ++i)
          if (m->row_ids[i] >= p->n_owned && m->row_sizes[i] > 0)
            p->cheb_fusable_multirank = false;
      }
      // halo overlap (hx_internal.h): which ghost rows the cell kernel pushes itself, which ones the closing kernel does
      if (p->nranks > 1)
        {
          p->overlap_x_ok = true;
          for (uint32_t e = 0; e < p->nnz; ++e)
            if (m->col_ids[e] >= p->n_owned)
              p->overlap_x_ok = false; // a hanging-node / periodic fill reads a ghost row: it must follow the unpack
          std::vector<uint32_t> unpack(p->n_ghost), rest;
          p->n_push_direct = 0;
          for (uint32_t k = 0; k < p->n_ghost; ++k)
            {
              const uint32_t j = m->halo.ghost_local_ids[k], r = p->n_owned + j;
              unpack[k]        = j | (rowinfo[r] == 0xFFFFFFFEu ? 0x80000000u : 0u);
              const bool direct = rowinfo[r] == 0xFFFFFFFFu && last_entry[r] != 0xffffffffu;
              if (direct)
                {
                  dest[last_entry[r]] |= HX_DEST_PUSH;
                  p->n_push_direct++;
                }
              else
                rest.push_back(k);
            }
          p->overlap_y_ok = true;
          p->n_push_rest  = (uint32_t)rest.size();
          HX_TRY(p->d_unpack_ids.upload(unpack));
          HX_TRY(p->d_push_rest.upload(rest));
          HX_TRY(p->d_x_ready.alloc(1));
          HX_CUDA(cudaMemset(p->d_x_ready.p, 0, sizeof(uint32_t)));
        }
      std::vector<uint32_t> untouched;
      for (uint32_t r = 0; r < p->n_local; ++r)
        if (last[r] == 0xffffffffu && shared.find(r) == shared.end())
          untouched.push_back(r);
      p->n_untouched = (uint32_t)untouched.size();
      HX_TRY(p->d_untouched.upload(untouched));
      HX_TRY(p->d_order.upload(p->h_order));
      HX_TRY(p->d_wait_off.upload(p->h_wait_off));
      HX_TRY(p->d_wait_list.upload(p->h_wait_list));
      const size_t nflags = (size_t)std::max(p->C, 1u) * ((p->max_block + 7) / 8);
      HX_TRY(p->d_flags.alloc(nflags));
      HX_CUDA(cudaMemset(p->d_flags.p, 0, nflags * sizeof(uint32_t)));
      HX_TRY(p->d_counters.alloc(2));
      HX_CUDA(cudaMemset(p->d_counters.p, 0, 2 * sizeof(uint32_t)));
      HX_TRY(p->d_clk.alloc(2));
      HX_CUDA(cudaMemset(p->d_clk.p, 0, 2 * sizeof(unsigned long long)));
      p->epoch = 0;
      int dev = 0;
      HX_CUDA(cudaGetDevice(&dev));
      HX_CUDA(cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, dev));
      if (const char *m = getenv("HXB200_SCATTER"))
        p->scatter_mode = (strcmp(m, "coloured") == 0) ? 1 : 0;
    }
    HX_TRY(p->d_dest.upload(dest));

    HX_TRY(p->halo.init(m->halo, p->max_block));
    HX_CUDA(cudaEventCreate(&p->ev0));
    HX_CUDA(cudaEventCreate(&p->ev1));
    return HX_OK;
  }

  // ---------------------------------------------------------------------------------------------
  static int
  cellop_apply(hx_op *op, double *X, double *Y, uint32_t B, int ugx, int ugy, const FuseArgs *fuse = nullptr,
               bool *fused_applied = nullptr, bool y_constrained_rows_dead = false)
  {
    hx_plan *p = op->plan;
    p->mark("apply:begin");
    // halo exchange overlapped with the cell kernel (north_star: "NCCL/NVLink halo exchange overlapped with interior-cell
    // compute"; the reference splits its exchange into Begin / End for the same purpose, MPICommunicatorP2P.t.cpp:89-273,
    // 288-470): the send side of the X update runs here, its receive side inside the cell kernel; the send side of the Y
    // accumulation runs inside the cell kernel, its receive side after it.  Needs the peer-memory transport, the mesh's own
    // constraint set on both sides and the ordered kernel; HXB200_HALO_OVERLAP=0 keeps the serial exchange (same results,
    // bit for bit: the processing order does not depend on the mode).
    bool overlap = false;
    if (p->nranks > 1 && p->scatter_mode == 0 && op->x_set == 0 && op->y_set == 0 && p->C > 0)
      {
        HX_CHECK(p->comm != nullptr, HX_ERR_COMM, "nranks > 1 but no communicator attached (hx_plan_attach_comm)");
        HX_TRY(halo_pick_transport(p, p->halo)); // collective
        const char *e = getenv("HXB200_HALO_OVERLAP");
        overlap       = peer_overlap_available(p->halo) && p->overlap_y_ok && !(e && e[0] == '0');
      }
    if (overlap && op->has_nl && op->nl_reads_ghost < 0)
      {
        // the projector pre-pass (nl_phase_a) runs before the cell kernel and reads the X rows of the projector cells: if one
        // of them reads ghost rows, the ghosts must be in place before it - the unpack cannot wait for the cell kernel
        op->nl_reads_ghost = 0;
        for (uint32_t c = 0; c < p->C; ++c)
          if (op->h_ncp[c] > 0 && p->h_boundary[c])
            op->nl_reads_ghost = 1;
      }
    const bool unpack_in_kernel = overlap && ugx && p->overlap_x_ok && !(op->has_nl && op->nl_reads_ghost == 1);
    uint32_t   seqU             = 0;
    if (ugx)
      {
        if (unpack_in_kernel)
          HX_TRY(peer_push_update(p, p->halo, X, B, &seqU));
        else
          HX_TRY(halo_update(p, p->halo, X, B));
      }
    p->mark("x-halo");
    HX_TRY(launch_p2c(p, X, B, op->x_set));
    if (p->scatter_mode == 1)
      HX_CUDA(cudaMemsetAsync(Y, 0, (size_t)p->n_local * B * sizeof(double), p->stream));
    else // first touchers store instead of add: only rows no cell writes need clearing
      HX_TRY(launch_zero_rows(p, Y, B, p->d_untouched.p, p->n_untouched));
    p->mark("p2c+zero");
    if (op->has_nl)
      {
        HX_TRY(launch_nl_phase_a(op, X, B));
        p->mark("nl-phase-a");
        if (p->nranks > 1)
          {
            // applyAllReduceOnCconjtransX + applyVOnCconjtransX
            HX_TRY(halo_accumulate_update(p, op->phalo, op->d_cx.p, B));
            HX_TRY(launch_row_scale(p, op->d_v.p, op->d_cx.p, op->d_cx.p, B, op->n_proj_local));
            p->mark("nl-halo");
          }
      }
    HaloK hk;
    if (overlap)
      HX_TRY(peer_overlap_args(p, p->halo, X, B, unpack_in_kernel, seqU, &hk));
    HX_TRY(launch_cell_apply(op, X, Y, B, fuse, fused_applied, overlap ? &hk : nullptr));
    p->mark("cell-kernel");
    HX_TRY(launch_shared_reduce(p, Y, B));
    // (a caller that never reads the constrained rows of Y - the fused filter's scratch on a single rank, where no halo
    // accumulation follows - saves their zeroing)
    HX_TRY(launch_c2p(p, Y, B, op->y_set, !(y_constrained_rows_dead && p->nranks == 1)));
    p->mark("shared+c2p");
    if (overlap)
      HX_TRY(peer_finish_accumulate(p, p->halo, Y, B, hk.seqA));
    else
      HX_TRY(halo_accumulate(p, p->halo, Y, B));
    if (ugy)
      HX_TRY(halo_update(p, p->halo, Y, B));
    p->mark("y-halo");
    return HX_OK;
  }

  static int
  diagop_apply(hx_op *op, double *X, double *Y, uint32_t B, int ugx, int ugy)
  {
    hx_plan *p = op->plan;
    if (op->variant == HX_DIAG_JACOBI)
      {
        // PreconditionerJacobi::apply (src/linearAlgebra/PreconditionerJacobi.t.cpp:52-82): no constraints involved
        if (ugx)
          HX_TRY(halo_update(p, p->halo, X, B));
        HX_TRY(launch_row_scale(p, op->d_diag.p, X, Y, B, p->n_local));
        if (ugy)
          HX_TRY(halo_update(p, p->halo, Y, B));
        return HX_OK;
      }
    if (op->variant == HX_DIAG_OEFE_GLOBAL)
      {
        // OrthoEFEOverlapInverseOpContextGLL::apply (src/basis/OrthoEFEOverlapInverseOpContextGLL.t.cpp:1182-1282): diagonal
        // part on every local row, then the dense block over ALL enrichment functions of the system: every rank places its
        // owned enrichment rows at their global offset, the vector is summed over the ranks (the reference's MPI_Allreduce
        // of nE_global x B, :1228-1234; ncclAllReduce here), and each rank takes its own rows of block . Xenr
        if (ugx)
          HX_TRY(halo_update(p, p->halo, X, B));
        HX_TRY(launch_p2c(p, X, B));
        HX_TRY(launch_row_scale(p, op->d_diag.p, X, Y, B, p->n_local));
        if (op->nE_global)
          {
            HX_TRY(p->ensure_small((size_t)op->nE_global * B));
            double *xg = p->d_small.p;
            HX_CUDA(cudaMemsetAsync(xg, 0, (size_t)op->nE_global * B * sizeof(double), p->stream));
            if (op->nE)
              HX_CUDA(cudaMemcpyAsync(xg + (size_t)op->enr_offset * B, X + (size_t)p->n_owned_classical * B,
                                      (size_t)op->nE * B * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
            if (p->nranks > 1)
              {
                HX_CHECK(p->comm != nullptr, HX_ERR_COMM, "nranks > 1 but no communicator attached (hx_plan_attach_comm)");
                HX_TRY(comm_allreduce_sum(p->comm, p->stream, xg, (size_t)op->nE_global * B));
              }
            HX_TRY(launch_enr_block_global(p, op->d_enr_block.p, op->nE_global, op->enr_offset, op->nE, xg,
                                           Y + (size_t)p->n_owned_classical * B, B));
          }
        HX_TRY(halo_update(p, p->halo, Y, B));
        HX_TRY(launch_c2p(p, Y, B));
        if (ugy)
          HX_TRY(halo_update(p, p->halo, Y, B));
        return HX_OK;
      }
    if (op->variant == HX_DIAG_OEFE_MASS)
      ugx = ugy = 0;
    if (ugx)
      HX_TRY(halo_update(p, p->halo, X, B));
    HX_TRY(launch_p2c(p, X, B));
    HX_TRY(launch_row_scale(p, op->d_diag.p, X, Y, B, p->n_local));
    if (op->variant != HX_DIAG_CFE && op->variant != HX_DIAG_JACOBI)
      {
        const size_t o = (size_t)p->n_owned_classical * B;
        HX_TRY(launch_enr_block(p, op->d_enr_block.p, op->nE, X + o, Y + o, B));
        HX_TRY(halo_update(p, p->halo, Y, B));
      }
    HX_TRY(launch_c2p(p, Y, B));
    if (ugy)
      HX_TRY(halo_update(p, p->halo, Y, B));
    return HX_OK;
  }

  int
  op_apply(hx_op *op, double *X, double *Y, uint32_t B, int ugx, int ugy)
  {
    HX_CHECK(op && X && Y, HX_ERR_INVALID, "null argument");
    HX_CHECK(B >= 1 && B <= op->plan->max_block, HX_ERR_INVALID, "B = %u outside [1, max_block = %u]", B,
             op->plan->max_block);
    HX_CHECK(X != Y, HX_ERR_INVALID, "X and Y must not alias");
    return op->kind == HX_OP_CELL ? cellop_apply(op, X, Y, B, ugx, ugy) : diagop_apply(op, X, Y, B, ugx, ugy);
  }

  __global__ void
  copy_cols_kernel(const double *src, uint32_t ldsrc, uint32_t c0s, double *dst, uint32_t lddst, uint32_t c0d,
                   uint32_t ncols, size_t nrows)
  {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * ncols)
      return;
    const size_t r = i / ncols;
    const uint32_t c = (uint32_t)(i % ncols);
    dst[r * lddst + c0d + c] = src[r * ldsrc + c0s + c];
  }
  int
  copy_cols(hx_plan *p, const double *src, uint32_t ldsrc, uint32_t c0s, double *dst, uint32_t lddst, uint32_t c0d,
            uint32_t ncols, size_t nrows)
  {
    const size_t tot = nrows * ncols;
    if (!tot)
      return HX_OK;
    copy_cols_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, p->stream>>>(src, ldsrc, c0s, dst, lddst, c0d, ncols, nrows);
    p->launches++;
    HX_CUDA(cudaGetLastError());
    return HX_OK;
  }
} // namespace hx

using namespace hx;

hx::ConstraintView
hx_plan::constraint_view(uint32_t set) const
{
  if (set > 0 && set <= extra_constraints.size())
    return extra_constraints[set - 1]->view();
  hx::ConstraintView v;
  v.nR = nR, v.nPar = nPar, v.max_row = max_row, v.max_child = max_child;
  v.row_ids = d_row_ids.p, v.row_sizes = d_row_sizes.p, v.row_offsets = d_row_offsets.p, v.col_ids = d_col_ids.p;
  v.col_vals = d_col_vals.p, v.inhom = d_inhom.p;
  v.par_ids = d_par_ids.p, v.par_off = d_par_off.p, v.par_child = d_par_child.p, v.par_w = d_par_w.p;
  return v;
}

hx_plan::~hx_plan()
{
  for (auto *c : extra_constraints)
    delete c;
  for (auto &m : trace_marks)
    cudaEventDestroy(m.second);
  for (auto e : trace_pool)
    cudaEventDestroy(e);
  for (auto *s : scratch)
    delete s;
  if (h_pinned)
    cudaFreeHost(h_pinned);
  if (ev0)
    cudaEventDestroy(ev0);
  if (ev1)
    cudaEventDestroy(ev1);
  for (auto e : ev_pool)
    cudaEventDestroy(e);
  if (comm)
    comm_destroy(comm);
  if (dense)
    dense_destroy(dense);
  for (auto e : pipe_ev)
    if (e)
      cudaEventDestroy(e);
  if (copy_in)
    cudaStreamDestroy(copy_in);
  if (copy_out)
    cudaStreamDestroy(copy_out);
  if (own_stream && stream)
    cudaStreamDestroy(stream);
}

int
hx_plan::get_scratch(size_t idx, double **out, uint32_t cols)
{
  // n_local x cols doubles (cols = 0: max_block), grow-only: a 60 M-DoF plan with max_block = 1024 filters 64 columns at
  // a time and must not pay 60 GB per scratch block
  while (scratch.size() <= idx)
    scratch.push_back(new DevBuf<double>());
  const size_t need = (size_t)n_local * (cols ? cols : max_block);
  if (scratch[idx]->n < need)
    HX_TRY(scratch[idx]->alloc(need));
  *out = scratch[idx]->p;
  return HX_OK;
}
void
hx_plan::mark(const char *name)
{
  if (!trace)
    return;
  cudaEvent_t e;
  if (trace_pool.empty())
    {
      if (cudaEventCreate(&e) != cudaSuccess)
        return;
    }
  else
    {
      e = trace_pool.back();
      trace_pool.pop_back();
    }
  cudaEventRecord(e, stream);
  trace_marks.push_back({name, e});
}

int
hx_plan::ensure_small(size_t doubles)
{
  if (d_small.n < doubles)
    HX_TRY(d_small.alloc(doubles));
  return HX_OK;
}
int
hx_plan::ensure_pinned(size_t bytes)
{
  if (h_pinned_bytes < bytes)
    {
      if (h_pinned)
        cudaFreeHost(h_pinned);
      h_pinned       = nullptr;
      h_pinned_bytes = 0;
      HX_CUDA(cudaMallocHost((void **)&h_pinned, bytes));
      h_pinned_bytes = bytes;
    }
  return HX_OK;
}

extern "C"
{
  const char *
  hx_last_error(void)
  {
    return g_err;
  }
  int
  hx_version(void)
  {
    return 100;
  }

  int
  hx_device_count(int *n)
  {
    HX_CUDA(cudaGetDeviceCount(n));
    return HX_OK;
  }
  int
  hx_set_device(int d)
  {
    HX_CUDA(cudaSetDevice(d));
    return HX_OK;
  }
  int
  hx_device_alloc(void **ptr, size_t bytes)
  {
    HX_CUDA(cudaMalloc(ptr, bytes ? bytes : 8));
    return HX_OK;
  }
  int
  hx_device_free(void *ptr)
  {
    HX_CUDA(cudaFree(ptr));
    return HX_OK;
  }
  int
  hx_host_alloc_pinned(void **ptr, size_t bytes)
  {
    HX_CUDA(cudaMallocHost(ptr, bytes ? bytes : 8));
    return HX_OK;
  }
  int
  hx_host_free_pinned(void *ptr)
  {
    HX_CUDA(cudaFreeHost(ptr));
    return HX_OK;
  }
  int
  hx_memcpy_h2d(void *d, const void *s, size_t bytes)
  {
    // plan streams are non-blocking (no implicit ordering with the legacy stream these helpers use):
    // make the helpers device-synchronous on both sides so a harness cannot race a plan's kernels
    HX_CUDA(cudaDeviceSynchronize());
    HX_CUDA(cudaMemcpy(d, s, bytes, cudaMemcpyHostToDevice));
    HX_CUDA(cudaDeviceSynchronize());
    return HX_OK;
  }
  int
  hx_memcpy_d2h(void *d, const void *s, size_t bytes)
  {
    HX_CUDA(cudaDeviceSynchronize());
    HX_CUDA(cudaMemcpy(d, s, bytes, cudaMemcpyDeviceToHost));
    return HX_OK;
  }
  int
  hx_memset_zero(void *d, size_t bytes)
  {
    HX_CUDA(cudaDeviceSynchronize());
    HX_CUDA(cudaMemset(d, 0, bytes));
    HX_CUDA(cudaDeviceSynchronize());
    return HX_OK;
  }

  int
  hx_plan_create(hx_plan **plan, const hx_mesh_desc *mesh, void *stream)
  {
    HX_CHECK(plan && mesh, HX_ERR_INVALID, "null argument");
    HX_CHECK(mesh->struct_size == sizeof(hx_mesh_desc), HX_ERR_INVALID, "hx_mesh_desc size mismatch (%u vs %zu)",
             mesh->struct_size, sizeof(hx_mesh_desc));
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      {
        set_error("no CUDA device available: libhxb200 has no CPU fallback");
        return HX_ERR_CUDA;
      }
    hx_plan *p = new (std::nothrow) hx_plan();
    HX_CHECK(p, HX_ERR_NOMEM, "out of host memory");
    if (stream)
      p->stream = (cudaStream_t)stream;
    else
      {
        cudaError_t e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess)
          {
            set_error("cudaStreamCreate: %s", cudaGetErrorString(e));
            delete p;
            return HX_ERR_CUDA;
          }
        p->own_stream = true;
      }
    int r = build_plan(p, mesh);
    if (r == HX_OK && cudaDeviceSynchronize() != cudaSuccess) // pageable uploads: DMA must have landed
      {
        set_error("cudaDeviceSynchronize failed after plan upload");
        r = HX_ERR_CUDA;
      }
    if (r != HX_OK)
      {
        delete p;
        return r;
      }
    *plan = p;
    return HX_OK;
  }

  int
  hx_plan_destroy(hx_plan *plan)
  {
    if (plan)
      {
        cudaStreamSynchronize(plan->stream);
        halo_peer_release(plan, plan->halo);
        delete plan;
      }
    return HX_OK;
  }

  int
  hx_plan_synchronize(hx_plan *plan)
  {
    HX_CHECK(plan, HX_ERR_INVALID, "null plan");
    return plan_sync(plan);
  }

  // MultiVector::globalSize(): the number of locally owned rows summed over the ranks of the plan's communicator
  int
  hx_plan_global_size(hx_plan *plan, uint64_t *n)
  {
    HX_CHECK(plan && n, HX_ERR_INVALID, "null argument");
    if (plan->n_global == 0)
      {
        double v = (double)plan->n_owned;
        if (plan->nranks > 1)
          {
            HX_CHECK(plan->comm != nullptr, HX_ERR_COMM, "nranks > 1 but no communicator attached (hx_plan_attach_comm)");
            HX_TRY(plan->ensure_small(1));
            HX_CUDA(cudaMemcpyAsync(plan->d_small.p, &v, sizeof(double), cudaMemcpyHostToDevice, plan->stream));
            HX_TRY(comm_allreduce_sum(plan->comm, plan->stream, plan->d_small.p, 1));
            HX_CUDA(cudaMemcpyAsync(&v, plan->d_small.p, sizeof(double), cudaMemcpyDeviceToHost, plan->stream));
            HX_TRY(plan_sync(plan));
          }
        plan->n_global = (uint64_t)(v + 0.5);
      }
    *n = plan->n_global;
    return HX_OK;
  }

  int
  hx_plan_halo_transport(hx_plan *plan, int *transport)
  {
    HX_CHECK(plan && transport, HX_ERR_INVALID, "null argument");
    *transport = plan->halo_transport;
    return HX_OK;
  }

  int
  hx_comm_unique_id(char id[128])
  {
    return comm_unique_id(id);
  }
  int
  hx_plan_attach_comm(hx_plan *plan, const char id[128])
  {
    HX_CHECK(plan, HX_ERR_INVALID, "null plan");
    if (plan->comm)
      {
        comm_destroy(plan->comm);
        plan->comm = nullptr;
      }
    return comm_create(&plan->comm, id, plan->nranks, plan->rank);
  }

  int
  hx_plan_set_scatter_mode(hx_plan *plan, int mode)
  {
    HX_CHECK(plan, HX_ERR_INVALID, "null plan");
    HX_CHECK(mode == 0 || mode == 1, HX_ERR_INVALID, "scatter mode must be 0 (ordered) or 1 (coloured)");
    plan->scatter_mode = mode;
    return HX_OK;
  }
  int
  hx_plan_get_processing_order(hx_plan *plan, uint32_t *order)
  {
    HX_CHECK(plan && order, HX_ERR_INVALID, "null argument");
    memcpy(order, plan->h_order.data(), sizeof(uint32_t) * plan->C);
    return HX_OK;
  }
  int
  hx_plan_get_wait_lists(hx_plan *plan, uint32_t *nnz, uint32_t *offsets, uint32_t *preds)
  {
    HX_CHECK(plan && nnz, HX_ERR_INVALID, "null argument");
    *nnz = (uint32_t)plan->h_wait_list.size();
    if (offsets)
      memcpy(offsets, plan->h_wait_off.data(), sizeof(uint32_t) * (plan->C + 1));
    if (preds)
      memcpy(preds, plan->h_wait_list.data(), sizeof(uint32_t) * plan->h_wait_list.size());
    return HX_OK;
  }

  int
  hx_programmatic_launch_enabled(void)
  {
    return hx::pdl_enabled() ? 1 : 0;
  }

  int
  hx_plan_num_colours(hx_plan *plan, uint32_t *n)
  {
    HX_CHECK(plan && n, HX_ERR_INVALID, "null argument");
    *n = plan->n_colours;
    return HX_OK;
  }
  int
  hx_plan_get_cell_colours(hx_plan *plan, uint32_t *colour)
  {
    HX_CHECK(plan && colour, HX_ERR_INVALID, "null argument");
    memcpy(colour, plan->h_colour.data(), sizeof(uint32_t) * plan->C);
    return HX_OK;
  }
  int
  hx_plan_get_fusable_rows(hx_plan *plan, uint32_t *n_fusable, uint32_t *n_other_owned)
  {
    HX_CHECK(plan && n_fusable && n_other_owned, HX_ERR_INVALID, "null argument");
    *n_fusable     = plan->n_fusable;
    *n_other_owned = plan->n_nonfuse;
    return HX_OK;
  }
  int
  hx_plan_get_c2p_transpose(hx_plan *plan, uint32_t *n_parents, uint32_t *parent_ids, uint32_t *offsets,
                            uint32_t *child_rows, double *weights)
  {
    HX_CHECK(plan && n_parents, HX_ERR_INVALID, "null argument");
    *n_parents = plan->nPar;
    if (parent_ids)
      memcpy(parent_ids, plan->h_par_ids.data(), sizeof(uint32_t) * plan->nPar);
    if (offsets)
      memcpy(offsets, plan->h_par_off.data(), sizeof(uint32_t) * (plan->nPar + 1));
    if (child_rows)
      memcpy(child_rows, plan->h_par_child.data(), sizeof(uint32_t) * plan->h_par_child.size());
    if (weights)
      memcpy(weights, plan->h_par_w.data(), sizeof(double) * plan->h_par_w.size());
    return HX_OK;
  }

  int
  hx_update_ghost_values(hx_plan *plan, double *X, uint32_t B)
  {
    HX_CHECK_B(plan, B);
    return halo_update(plan, plan->halo, X, B);
  }
  int
  hx_accumulate_add_locally_owned(hx_plan *plan, double *Y, uint32_t B)
  {
    HX_CHECK_B(plan, B);
    return halo_accumulate(plan, plan->halo, Y, B);
  }
  int
  hx_distribute_parent_to_child(hx_plan *plan, double *X, uint32_t B)
  {
    HX_CHECK_B(plan, B);
    return launch_p2c(plan, X, B);
  }
  int
  hx_distribute_child_to_parent(hx_plan *plan, double *Y, uint32_t B)
  {
    HX_CHECK_B(plan, B);
    return launch_c2p(plan, Y, B);
  }
  int
  hx_set_constrained_nodes_to_zero(hx_plan *plan, double *Y, uint32_t B)
  {
    HX_CHECK_B(plan, B);
    return launch_zero_constrained(plan, Y, B);
  }

  int
  hx_cellop_create(hx_plan *plan, hx_op **op)
  {
    HX_CHECK(plan && op, HX_ERR_INVALID, "null argument");
    hx_op *o = new (std::nothrow) hx_op();
    HX_CHECK(o, HX_ERR_NOMEM, "out of host memory");
    o->plan = plan;
    o->kind = HX_OP_CELL;
    *op     = o;
    return HX_OK;
  }

  int
  hx_plan_add_constraints(hx_plan *plan, uint32_t n_rows, const uint32_t *row_ids, const uint32_t *row_sizes,
                          const uint32_t *row_offsets, const uint32_t *col_ids, const double *col_vals, const double *inhom,
                          uint32_t *set_id)
  {
    HX_CHECK(plan && set_id, HX_ERR_INVALID, "null argument");
    HX_CHECK(n_rows == 0 || (row_ids && row_sizes && row_offsets && inhom), HX_ERR_INVALID, "null constraint arrays");
    std::unique_ptr<ConstraintSet> c(new (std::nothrow) ConstraintSet());
    HX_CHECK(c, HX_ERR_NOMEM, "out of host memory");
    c->nR = n_rows;
    std::vector<char> constrained(plan->n_local, 0);
    for (uint32_t i = 0; i < n_rows; ++i)
      {
        HX_CHECK(row_ids[i] < plan->n_local, HX_ERR_INVALID, "constraint row id out of range");
        HX_CHECK(!constrained[row_ids[i]], HX_ERR_INVALID, "duplicate constraint row %u", row_ids[i]);
        constrained[row_ids[i]] = 1;
        c->nnz                  = std::max(c->nnz, row_offsets[i] + row_sizes[i]);
        c->max_row              = std::max(c->max_row, row_sizes[i]);
      }
    std::map<uint32_t, std::vector<std::pair<uint32_t, double>>> par;
    for (uint32_t i = 0; i < n_rows; ++i)
      for (uint32_t j = 0; j < row_sizes[i]; ++j)
        {
          const uint32_t col = col_ids[row_offsets[i] + j];
          HX_CHECK(col < plan->n_local, HX_ERR_INVALID, "constraint column id out of range");
          HX_CHECK(!constrained[col], HX_ERR_UNSUPPORTED, "constraint chain: row %u depends on constrained row %u", row_ids[i],
                   col);
          par[col].push_back({row_ids[i], col_vals[row_offsets[i] + j]});
        }
    std::vector<uint32_t> par_ids, par_off(1, 0), par_child;
    std::vector<double>   par_w;
    for (auto &kv : par)
      {
        par_ids.push_back(kv.first);
        for (auto &e : kv.second)
          {
            par_child.push_back(e.first);
            par_w.push_back(e.second);
          }
        par_off.push_back((uint32_t)par_child.size());
        c->max_child = std::max(c->max_child, (uint32_t)kv.second.size());
      }
    c->nPar = (uint32_t)par_ids.size();
    HX_TRY(c->d_row_ids.upload(row_ids, n_rows));
    HX_TRY(c->d_row_sizes.upload(row_sizes, n_rows));
    HX_TRY(c->d_row_offsets.upload(row_offsets, n_rows));
    HX_TRY(c->d_col_ids.upload(col_ids, c->nnz));
    HX_TRY(c->d_col_vals.upload(col_vals, c->nnz));
    HX_TRY(c->d_inhom.upload(inhom, n_rows));
    HX_TRY(c->d_par_ids.upload(par_ids));
    HX_TRY(c->d_par_off.upload(par_off));
    HX_TRY(c->d_par_child.upload(par_child));
    HX_TRY(c->d_par_w.upload(par_w));
    HX_CUDA(cudaDeviceSynchronize());
    plan->extra_constraints.push_back(c.release());
    *set_id = (uint32_t)plan->extra_constraints.size();
    return HX_OK;
  }

  int
  hx_cellop_set_constraint_sets(hx_op *op, uint32_t x_set, uint32_t y_set)
  {
    HX_CHECK(op && op->kind == HX_OP_CELL, HX_ERR_INVALID, "bad argument");
    const uint32_t n = (uint32_t)op->plan->extra_constraints.size();
    HX_CHECK(x_set <= n && y_set <= n, HX_ERR_INVALID, "unknown constraint set");
    op->x_set = x_set;
    op->y_set = y_set;
    return HX_OK;
  }

  int
  hx_cellop_set_matrix_sharing(hx_op *op, int enable)
  {
    HX_CHECK(op && op->kind == HX_OP_CELL, HX_ERR_INVALID, "not a cell operator");
    op->share_identical = enable != 0;
    return HX_OK;
  }
  int
  hx_cellop_num_unique_matrices(hx_op *op, uint32_t *n)
  {
    HX_CHECK(op && n && op->kind == HX_OP_CELL, HX_ERR_INVALID, "not a cell operator");
    HX_CHECK(op->have_matrices, HX_ERR_INVALID, "cell operator has no matrices (call hx_cellop_set_matrices)");
    *n = op->n_unique;
    return HX_OK;
  }

  int
  hx_cellop_set_nonlocal(hx_op *op, const hx_nonlocal_desc *nl)
  {
    HX_CHECK(op && nl && op->kind == HX_OP_CELL, HX_ERR_INVALID, "bad argument");
    HX_CHECK(nl->struct_size == sizeof(hx_nonlocal_desc), HX_ERR_INVALID, "hx_nonlocal_desc size mismatch");
    hx_plan *p       = op->plan;
    op->have_matrices = false; // the packed layout carries the projector columns: matrices must be (re)set
    op->h_ncp.assign(nl->num_cell_proj, nl->num_cell_proj + p->C);
    for (uint32_t c = 0; c < p->C; ++c)
      HX_CHECK(op->h_ncp[c] <= 64, HX_ERR_UNSUPPORTED, "cell %u couples to %u projectors (the projector pre-pass holds at most 64 per cell)", c,
               op->h_ncp[c]);
    op->sum_proj = 0;
    op->h_c_off.assign(p->C + 1, 0);
    op->h_nl_cells.clear();
    std::vector<unsigned long long> coff(p->C + 1, 0);
    for (uint32_t c = 0; c < p->C; ++c)
      {
        op->sum_proj += op->h_ncp[c];
        coff[c + 1] = coff[c] + (unsigned long long)op->h_ncp[c] * p->h_ncd[c];
        if (op->h_ncp[c])
          op->h_nl_cells.push_back(c);
      }
    op->n_proj_local = nl->proj_halo.n_owned + nl->proj_halo.n_ghost;
    op->h_pids.assign(nl->cell_proj_local_ids, nl->cell_proj_local_ids + op->sum_proj);
    for (uint32_t i = 0; i < op->sum_proj; ++i)
      HX_CHECK(op->h_pids[i] < op->n_proj_local, HX_ERR_INVALID, "projector id out of range");
    HX_TRY(op->d_pids.upload(op->h_pids));
    HX_TRY(op->d_nl_cells.upload(op->h_nl_cells));
    HX_TRY(op->d_cell_c.upload(nl->cell_c, (size_t)coff[p->C]));
    HX_TRY(op->d_c_off.upload(coff.data(), coff.size()));
    HX_TRY(op->d_v.upload(nl->v, op->n_proj_local));
    HX_TRY(op->d_cx.alloc((size_t)std::max(op->n_proj_local, 1u) * p->max_block));
    HX_TRY(op->d_cx_stage.alloc((size_t)std::max(op->sum_proj, 1u) * p->max_block));
    HX_CUDA(cudaMemset(op->d_cx.p, 0, op->d_cx.n * sizeof(double)));
    // projector row -> staging slots (slot = position in cell_proj_local_ids, i.e. ascending cell order)
    std::vector<uint32_t> cnt(op->n_proj_local + 1, 0), slots(op->sum_proj);
    for (uint32_t i = 0; i < op->sum_proj; ++i)
      cnt[op->h_pids[i] + 1]++;
    for (uint32_t r = 0; r < op->n_proj_local; ++r)
      cnt[r + 1] += cnt[r];
    {
      std::vector<uint32_t> fill(cnt.begin(), cnt.end() - 1);
      for (uint32_t i = 0; i < op->sum_proj; ++i)
        slots[fill[op->h_pids[i]]++] = i;
    }
    HX_TRY(op->d_pr_off.upload(cnt));
    HX_TRY(op->d_pr_slots.upload(slots));
    HX_TRY(op->phalo.init(nl->proj_halo, p->max_block));
    HX_CUDA(cudaDeviceSynchronize());
    op->has_nl = true;
    op->nl_reads_ghost = -1;
    return HX_OK;
  }

  int
  hx_cellop_set_matrices(hx_op *op, const double *cell_matrices, int on_device)
  {
    HX_CHECK(op && cell_matrices && op->kind == HX_OP_CELL, HX_ERR_INVALID, "bad argument");
    return pack_cell_matrices(op, cell_matrices, on_device);
  }

  int
  hx_diagop_create(hx_plan *plan, const double *diag, const double *enr_block, int variant, hx_op **op)
  {
    HX_CHECK(plan && diag && op, HX_ERR_INVALID, "null argument");
    HX_CHECK(variant >= HX_DIAG_CFE && variant <= HX_DIAG_JACOBI, HX_ERR_INVALID, "bad variant");
    hx_op *o = new (std::nothrow) hx_op();
    HX_CHECK(o, HX_ERR_NOMEM, "out of host memory");
    o->plan    = plan;
    o->kind    = HX_OP_DIAG;
    o->variant = variant;
    o->nE      = plan->n_owned - plan->n_owned_classical;
    int r      = o->d_diag.upload(diag, plan->n_local);
    if (r == HX_OK && variant != HX_DIAG_CFE && variant != HX_DIAG_JACOBI && o->nE)
      {
        if (!enr_block)
          {
            set_error("enrichment block required for %u owned enrichment rows", o->nE);
            r = HX_ERR_INVALID;
          }
        else
          r = o->d_enr_block.upload(enr_block, (size_t)o->nE * o->nE);
      }
    if (r == HX_OK && cudaDeviceSynchronize() != cudaSuccess)
      {
        set_error("cudaDeviceSynchronize failed after operator upload");
        r = HX_ERR_CUDA;
      }
    if (r != HX_OK)
      {
        delete o;
        return r;
      }
    *op = o;
    return HX_OK;
  }

  int
  hx_diagop_create_global_enrichment(hx_plan *plan, const double *diag_inv, const double *enr_block_global, uint32_t nE_global,
                                     uint32_t owned_enr_offset, hx_op **op)
  {
    HX_CHECK(plan && diag_inv && op, HX_ERR_INVALID, "null argument");
    const uint32_t nE = plan->n_owned - plan->n_owned_classical;
    HX_CHECK((uint64_t)owned_enr_offset + nE <= nE_global, HX_ERR_INVALID,
             "owned enrichment rows [%u, %u) outside the global enrichment range [0, %u)", owned_enr_offset, owned_enr_offset + nE,
             nE_global);
    HX_CHECK(nE_global == 0 || enr_block_global, HX_ERR_INVALID, "enrichment block required");
    hx_op *o = new (std::nothrow) hx_op();
    HX_CHECK(o, HX_ERR_NOMEM, "out of host memory");
    o->plan       = plan;
    o->kind       = HX_OP_DIAG;
    o->variant    = HX_DIAG_OEFE_GLOBAL;
    o->nE         = nE;
    o->nE_global  = nE_global;
    o->enr_offset = owned_enr_offset;
    int r         = o->d_diag.upload(diag_inv, plan->n_local);
    if (r == HX_OK && nE_global)
      r = o->d_enr_block.upload(enr_block_global, (size_t)nE_global * nE_global);
    if (r == HX_OK && cudaDeviceSynchronize() != cudaSuccess)
      {
        set_error("cudaDeviceSynchronize failed after operator upload");
        r = HX_ERR_CUDA;
      }
    if (r != HX_OK)
      {
        delete o;
        return r;
      }
    *op = o;
    return HX_OK;
  }

  int
  hx_op_destroy(hx_op *op)
  {
    if (op)
      {
        cudaStreamSynchronize(op->plan->stream);
        halo_peer_release(op->plan, op->phalo);
        delete op;
      }
    return HX_OK;
  }

  int
  hx_op_apply(hx_op *op, double *X, double *Y, uint32_t B, int ugx, int ugy)
  {
    return op_apply(op, X, Y, B, ugx, ugy);
  }

  int
  hx_op_apply_host(hx_op *op, double *Xh, double *Yh, uint32_t B, int ugx, int ugy)
  {
    HX_CHECK(op && Xh && Yh, HX_ERR_INVALID, "null argument");
    hx_plan *p = op->plan;
    HX_CHECK_B(p, B);
    double *dX, *dY;
    HX_TRY(p->get_scratch(4, &dX));
    HX_TRY(p->get_scratch(5, &dY));
    const size_t bytes = (size_t)p->n_local * B * sizeof(double);
    HX_CUDA(cudaMemcpyAsync(dX, Xh, bytes, cudaMemcpyHostToDevice, p->stream));
    HX_TRY(op_apply(op, dX, dY, B, ugx, ugy));
    HX_CUDA(cudaMemcpyAsync(Yh, dY, bytes, cudaMemcpyDeviceToHost, p->stream));
    // X is modified in place by the operator (OperatorContext.h:98-101): only the constrained rows (hanging-node
    // fill) and, after a ghost update, the ghost rows change - bring just those back
    const bool     ghosts = ugx && p->nranks > 1 && p->n_ghost > 0;
    const uint32_t nmod   = p->nR + (ghosts ? p->n_ghost : 0);
    if (nmod)
      {
        if (p->d_modrows.n != (size_t)p->nR + p->n_ghost)
          {
            std::vector<uint32_t> rows(p->h_row_ids);
            for (uint32_t g = 0; g < p->n_ghost; ++g)
              rows.push_back(p->n_owned + g);
            HX_TRY(p->d_modrows.upload(rows));
            p->h_modrows = rows;
            HX_CUDA(cudaDeviceSynchronize());
          }
        double *buf;
        HX_TRY(p->get_scratch(7, &buf));
        HX_TRY(launch_pack(p, dX, B, p->d_modrows.p, nmod, buf));
        HX_TRY(p->ensure_pinned((size_t)nmod * B * sizeof(double)));
        HX_CUDA(cudaMemcpyAsync(p->h_pinned, buf, (size_t)nmod * B * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
      }
    HX_TRY(plan_sync(p));
    for (uint32_t i = 0; i < nmod; ++i)
      memcpy(Xh + (size_t)p->h_modrows[i] * B, p->h_pinned + (size_t)i * B, (size_t)B * sizeof(double));
    return HX_OK;
  }

  int
  hx_distribute_parent_to_child_set(hx_plan *plan, uint32_t set, double *X, uint32_t B)
  {
    HX_CHECK_B(plan, B);
    HX_CHECK(set <= plan->extra_constraints.size(), HX_ERR_INVALID, "unknown constraint set");
    return launch_p2c(plan, X, B, set);
  }
  int
  hx_distribute_child_to_parent_set(hx_plan *plan, uint32_t set, double *Y, uint32_t B)
  {
    HX_CHECK_B(plan, B);
    HX_CHECK(set <= plan->extra_constraints.size(), HX_ERR_INVALID, "unknown constraint set");
    return launch_c2p(plan, Y, B, set);
  }

  // ------------------------------------------------------------------------------ CG solver ----
  // CGLinearSolver::solve (src/linearAlgebra/CGLinearSolver.t.cpp:68-300), scalar for scalar: per-column step
  // lengths, the residual is b - A x, convergence per column against max(absTol, |b| relTol) with the converged
  // column frozen in xConverged, divergence when a residual norm exceeds divTol.
  int
  hx_cg_solve(hx_op *A, hx_op *PC, const double *b, double *x, uint32_t B, uint32_t max_iter, double abs_tol,
              double rel_tol, double div_tol, uint32_t *iterations, int *status, double *residual_norms_host)
  {
    HX_CHECK(A && PC && b && x && iterations && status, HX_ERR_INVALID, "null argument");
    hx_plan *p = A->plan;
    HX_CHECK(p == PC->plan, HX_ERR_INVALID, "operators belong to different plans");
    HX_CHECK_B(p, B);
    HX_CHECK(B <= 256, HX_ERR_UNSUPPORTED, "hx_cg_solve supports B <= 256 right-hand sides per call");
    const size_t nloc = (size_t)p->n_local * B;
    double *     r, *w, *z, *pd, *xconv;
    HX_TRY(p->get_scratch(0, &r));
    HX_TRY(p->get_scratch(1, &w));
    HX_TRY(p->get_scratch(2, &z));
    HX_TRY(p->get_scratch(3, &pd));
    HX_TRY(p->get_scratch(6, &xconv));
    // per-column scalars stay on the device (one host synchronisation per iteration: the residual norms of the
    // convergence test); layout after the reduction scratch: ones | zdotr | pdotw | zdotr_new | rr | alpha | -alpha | beta.
    // With the Jacobi preconditioner on one rank the BLAS-1 work of an iteration is two fused passes (dots of z.r and
    // p.w; then x += alpha p, r -= alpha w, z = D^-1 r with the dots z.r and r.r) instead of ten launches.
    HX_TRY(p->ensure_small((size_t)1200 * B + 8 * (size_t)B));
    double *d_ones = p->d_small.p + (size_t)1200 * B, *d_zdotr = d_ones + B, *d_pdotw = d_zdotr + B, *d_zdotr_new = d_pdotw + B,
           *d_rr = d_zdotr_new + B, *d_alpha = d_rr + B, *d_nalpha = d_alpha + B, *d_beta = d_nalpha + B;
    HX_TRY(p->ensure_pinned(2 * (size_t)B * sizeof(double)));
    double *            h = p->h_pinned;
    std::vector<double> ones(B, 1.0), bnorm(B), rnorm(B);
    std::vector<char>   converged(B, 0);
    HX_CUDA(cudaMemcpyAsync(d_ones, ones.data(), B * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    std::vector<double> neg(B, -1.0);
    HX_CUDA(cudaMemcpyAsync(d_nalpha, neg.data(), B * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    HX_TRY(plan_sync(p));
    auto reduce = [&](const double *u, const double *v, double *out_dev) -> int {
      HX_TRY(launch_coldot(p, u, v, B, p->n_owned, out_dev));
      if (p->nranks > 1)
        HX_TRY(comm_allreduce_sum(p->comm, p->stream, out_dev, B));
      return HX_OK;
    };
    auto fetch = [&](const double *dev, double *out_host) -> int {
      HX_CUDA(cudaMemcpyAsync(h, dev, B * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
      HX_TRY(plan_sync(p));
      memcpy(out_host, h, B * sizeof(double));
      return HX_OK;
    };
    // out = 1*u + c[j]*v per column (linearAlgebra::add), over the owned rows; c lives on the device
    auto add = [&](const double *u, const double *c_dev, const double *v, double *out) -> int {
      return launch_axpby_blocked(p, p->n_owned, B, 1.0, d_ones, u, 1.0, c_dev, v, out);
    };
    HX_TRY(reduce(b, b, d_rr));
    HX_TRY(fetch(d_rr, bnorm.data()));
    for (uint32_t j = 0; j < B; ++j)
      bnorm[j] = sqrt(bnorm[j]);
    HX_CUDA(cudaMemcpyAsync(xconv, x, nloc * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    const bool fused_blas1 = p->nranks == 1 && PC->kind == HX_OP_DIAG && PC->variant == HX_DIAG_JACOBI && B <= 128 &&
                             !getenv("HXB200_CG_UNFUSED");
    int      err  = HX_CG_OTHER_ERROR; // until a column converges
    bool     diverged = false, all_conv = false;
    uint32_t iter = 0;
    for (; iter <= max_iter; ++iter)
      {
        if (iter == 0)
          {
            HX_TRY(op_apply(A, x, w, B, 1, 1));
            HX_TRY(add(b, d_nalpha, w, r)); // r = b - A x   (d_nalpha holds -1 here)
            HX_TRY(op_apply(PC, r, z, B, 0, 0));
            HX_CUDA(cudaMemcpyAsync(pd, z, nloc * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
          }
        else
          {
            HX_TRY(op_apply(A, pd, w, B, 1, 1));
            if (fused_blas1)
              {
                HX_TRY(launch_cg_dots2(p, z, r, pd, w, B, d_zdotr, d_pdotw, d_alpha, d_nalpha)); // alpha = z.r / p.w
                HX_TRY(launch_cg_update(p, x, pd, r, w, z, PC->d_diag.p, d_alpha, B, d_zdotr_new, d_rr, d_zdotr, d_beta));
                HX_TRY(add(z, d_beta, pd, pd)); // p = z + beta p
              }
            else
              {
                HX_TRY(reduce(z, r, d_zdotr));
                HX_TRY(reduce(pd, w, d_pdotw));
                HX_TRY(launch_col_divide(p, d_zdotr, d_pdotw, d_alpha, d_nalpha, B)); // alpha = z.r / p.w
                HX_TRY(add(x, d_alpha, pd, x));                                        // x += alpha p
                HX_TRY(add(r, d_nalpha, w, r));                                        // r -= alpha w
                HX_TRY(op_apply(PC, r, z, B, 0, 0));
                HX_TRY(reduce(z, r, d_zdotr_new));
                HX_TRY(launch_col_divide(p, d_zdotr_new, d_zdotr, d_beta, nullptr, B)); // beta = z.r (new) / z.r
                HX_TRY(add(z, d_beta, pd, pd));                                         // p = z + beta p
              }
          }
        if (!(fused_blas1 && iter > 0))
          HX_TRY(reduce(r, r, d_rr));
        HX_TRY(fetch(d_rr, rnorm.data()));
        for (uint32_t j = 0; j < B; ++j)
          {
            rnorm[j] = sqrt(rnorm[j]);
            if (rnorm[j] < std::max(abs_tol, bnorm[j] * rel_tol) && !converged[j])
              {
                err          = HX_CG_SUCCESS;
                converged[j] = 1;
                HX_TRY(copy_cols(p, x, B, j, xconv, B, j, 1, p->n_owned));
              }
            if (rnorm[j] > div_tol && !diverged)
              {
                err      = HX_CG_RESIDUAL_DIVERGENCE;
                diverged = true;
              }
          }
        all_conv = true;
        for (uint32_t j = 0; j < B; ++j)
          all_conv = all_conv && converged[j];
        if (diverged || all_conv)
          break;
      }
    // linearSolverFunction.setSolution(xConverged)
    HX_CUDA(cudaMemcpyAsync(x, xconv, nloc * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    HX_TRY(plan_sync(p));
    if (iter > max_iter)
      err = HX_CG_FAILED_TO_CONVERGE;
    *iterations = iter;
    *status     = err;
    if (residual_norms_host)
      memcpy(residual_norms_host, rnorm.data(), B * sizeof(double));
    return HX_OK;
  }

  // ------------------------------------------------------------------------------- filters ----
  // the rows of the fused filter's row list reordered as [rows without a child list | parent rows] (built on first use)
  static int
  build_split_row_list(hx_plan *p)
  {
    if (p->d_nonfuse_split.p || p->n_nonfuse == 0)
      return HX_OK;
    std::vector<uint32_t> rows(p->n_nonfuse), info(p->n_local), split;
    HX_TRY(plan_sync(p));
    HX_CUDA(cudaMemcpy(rows.data(), p->d_nonfuse_rows.p, rows.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    HX_CUDA(cudaMemcpy(info.data(), p->d_rowinfo.p, info.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    split.reserve(rows.size());
    for (uint32_t r : rows)
      if (info[r] >= 0xFFFFFFFEu) // free row without children, or constrained row
        split.push_back(r);
    p->n_nonfuse_plain = (uint32_t)split.size();
    for (uint32_t r : rows)
      if (info[r] < 0xFFFFFFFEu) // parent row: index into the parent-side CSR
        split.push_back(r);
    return p->d_nonfuse_split.upload(split);
  }

  int
  hx_chebyshev_filter(hx_op *A, hx_op *BInv, double *X, double *Y, uint32_t B, uint32_t degree, double a0, double a,
                      double b)
  {
    HX_CHECK(A && BInv && X && Y, HX_ERR_INVALID, "null argument");
    HX_CHECK(A->plan == BInv->plan, HX_ERR_INVALID, "operators belong to different plans");
    HX_CHECK(degree >= 1, HX_ERR_INVALID, "polynomial degree must be >= 1");
    HX_CHECK(X != Y, HX_ERR_INVALID, "X and Y must not alias (the recurrence ping-pongs between them)");
    hx_plan *p = A->plan;
    HX_CHECK_B(p, B);
    // ChebyshevFilter.t.cpp:62-70
    const double e      = 0.5 * (b - a);
    const double c      = 0.5 * (b + a);
    double       sigma  = e / (a0 - c);
    const double sigma1 = sigma;
    const double gamma  = 2.0 / sigma1;
    const size_t nown   = (size_t)p->n_owned * B;
    // the fused M^-1 needs no ghost update only if NO rank has a constrained ghost row with parents: the ranks must
    // agree (the unfused path contains an extra collective exchange), so the local flags are AND-ed once
    if (p->nranks > 1 && !p->cheb_fusable_agreed && BInv->kind == HX_OP_DIAG && BInv->variant != HX_DIAG_CFE)
      {
        HX_CHECK(p->comm != nullptr, HX_ERR_COMM, "nranks > 1 but no communicator attached (hx_plan_attach_comm)");
        std::vector<unsigned char> f1(4, 0), fall(4 * (size_t)p->nranks, 0);
        f1[0] = p->cheb_fusable_multirank ? 1 : 0;
        HX_TRY(comm_allgather_bytes(p->comm, p->stream, f1.data(), fall.data(), 4));
        for (int q = 0; q < p->nranks; ++q)
          if (!fall[4 * (size_t)q])
            p->cheb_fusable_multirank = false;
        p->cheb_fusable_agreed = true;
      }
    const bool   fused  = BInv->kind == HX_OP_DIAG && X != Y && BInv->variant != HX_DIAG_OEFE_GLOBAL &&
                       (BInv->variant == HX_DIAG_CFE || p->nranks == 1 || p->cheb_fusable_multirank);
    double *     s1, *s2 = nullptr;
    HX_TRY(p->get_scratch(0, &s1, B));
    if (!fused)
      HX_TRY(p->get_scratch(1, &s2, B));
    double *cur = X, *oth = Y; // cur = "eigenSubspaceGuess", oth = "filteredSubspace"
    // one degree of the mass-lumped path: s1 = A xc (updateGhostX), then out = ca*M^-1 s1 + cb*xc + cc*xp.  The
    // update of most rows happens inside the cell kernel's scatter (FuseArgs); the remaining owned rows (and all
    // rows, when the launched kernel variant cannot fuse) go through cheb_fused_kernel.
    const bool epilogue = fused && A->kind == HX_OP_CELL && p->scatter_mode == 0 && A->x_set == 0 && A->y_set == 0 &&
                          !getenv("HXB200_NO_EPILOGUE_FUSION");
    auto       degree_fused = [&](double *xc, const double *xp, double *out, double ca, double cb, double cc) -> int {
      bool applied = false;
      if (epilogue)
        {
          FuseArgs f;
          f.dinv = BInv->d_diag.p, f.xprev = xp, f.out = out, f.a = ca, f.b = cb, f.c = xp ? cc : 0.0;
          HX_CHECK(xc != s1 && out != xc, HX_ERR_INVALID, "aliasing in the fused Chebyshev step");
          // s1 is scratch: its constrained rows are either refilled by the hanging-node fill below or never read
          HX_TRY(cellop_apply(A, xc, s1, B, 1, 0, &f, &applied, true));
        }
      else
        HX_TRY(op_apply(A, xc, s1, B, 1, 0));
      if (!(p->cheb_fill_dead && p->nranks == 1))
        HX_TRY(launch_p2c(p, s1, B));
      int r;
      const char *split_env = getenv("HXB200_SPLIT_ROWLIST");
      const bool  split    = applied && !(split_env && split_env[0] == '0');
      if (split)
        HX_TRY(build_split_row_list(p));
      if (split && p->d_nonfuse_split.p && p->n_nonfuse_plain < p->n_nonfuse)
        {
          // the rows without a child list in one launch (no chain, 40 registers), the parent rows in a second one that
          // alone pays for the deep-chain variant.  Same kernels, same per-row work (HXB200_SPLIT_ROWLIST=0: one launch).
          r = launch_cheb_fused(p, BInv, s1, xc, xp, out, B, ca, cb, cc, true, p->d_nonfuse_split.p, p->n_nonfuse_plain, 1);
          if (r == HX_OK)
            r = launch_cheb_fused(p, BInv, s1, xc, xp, out, B, ca, cb, cc, true, p->d_nonfuse_split.p + p->n_nonfuse_plain,
                                  p->n_nonfuse - p->n_nonfuse_plain);
        }
      else
        r = applied ? launch_cheb_fused(p, BInv, s1, xc, xp, out, B, ca, cb, cc, true, p->d_nonfuse_rows.p, p->n_nonfuse) :
                      launch_cheb_fused(p, BInv, s1, xc, xp, out, B, ca, cb, cc);
      p->mark("cheb-rest");
      return r;
    };
    if (fused)
      HX_TRY(degree_fused(cur, nullptr, oth, sigma1 / e, -sigma1 / e * c, 0.0));
    else
      {
        HX_TRY(op_apply(A, cur, s1, B, 1, 0));
        HX_TRY(op_apply(BInv, s1, s2, B, 0, 0));
        HX_TRY(launch_axpby(p, nown, sigma1 / e, s2, -sigma1 / e * c, cur, oth));
      }
    for (uint32_t deg = 2; deg <= degree; ++deg)
      {
        const double sigma2 = 1.0 / (gamma - sigma);
        if (fused)
          HX_TRY(degree_fused(oth, cur, cur, 2.0 * sigma2 / e, -2.0 * sigma2 / e * c, -sigma * sigma2));
        else
          {
            HX_TRY(op_apply(A, oth, s1, B, 1, 0));
            HX_TRY(op_apply(BInv, s1, s2, B, 0, 0));
            HX_TRY(launch_axpby(p, nown, 2.0 * sigma2 / e, s2, -2.0 * sigma2 / e * c, oth, s1));
            HX_TRY(launch_axpby(p, nown, 1.0, s1, -sigma * sigma2, cur, cur));
          }
        std::swap(cur, oth);
        sigma = sigma2;
      }
    // result is in `oth`; the reference ends with eigenSubspaceGuess = filteredSubspace: both hold it
    HX_CUDA(cudaMemcpyAsync(cur, oth, (size_t)p->n_local * B * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    return HX_OK;
  }

  int
  hx_chebyshev_filter_host(hx_op *A, hx_op *BInv, double *Xh, double *Yh, uint32_t B, uint32_t degree, double a0,
                           double a, double b, int write_back_x)
  {
    HX_CHECK(A && BInv && Xh && Yh, HX_ERR_INVALID, "null argument");
    hx_plan *p = A->plan;
    HX_CHECK_B(p, B);
    double *dX, *dY;
    HX_TRY(p->get_scratch(4, &dX, B));
    HX_TRY(p->get_scratch(5, &dY, B));
    const size_t bytes = (size_t)p->n_local * B * sizeof(double);
    HX_CUDA(cudaMemcpyAsync(dX, Xh, bytes, cudaMemcpyHostToDevice, p->stream));
    HX_TRY(hx_chebyshev_filter(A, BInv, dX, dY, B, degree, a0, a, b));
    HX_CUDA(cudaMemcpyAsync(Yh, dY, bytes, cudaMemcpyDeviceToHost, p->stream));
    if (write_back_x)
      HX_CUDA(cudaMemcpyAsync(Xh, dX, bytes, cudaMemcpyDeviceToHost, p->stream));
    HX_TRY(plan_sync(p));
    return HX_OK;
  }

  // The column batches of ChebyshevFilteredEigenSolver::solve (src/linearAlgebra/ChebyshevFilteredEigenSolver.t.cpp:231-335:
  // the block of wavefunctions is filtered MAX_WAVEFN_BATCH_SIZE columns at a time) for wavefunctions that live in HOST
  // memory: batch k+1 is copied in and batch k-1 copied out while batch k is filtered - three streams, two device buffers
  // per direction - so the filter runs at its resident rate and PCIe is hidden behind it.  Xh[k] / Yh[k]: n_local x B,
  // contiguous, ideally pinned.
  int
  hx_chebyshev_filter_host_batches(hx_op *A, hx_op *BInv, const double *const *Xh, double *const *Yh, uint32_t n_batches,
                                   uint32_t B, uint32_t degree, double a0, double a, double b)
  {
    HX_CHECK(A && BInv && Xh && Yh, HX_ERR_INVALID, "null argument");
    hx_plan *p = A->plan;
    HX_CHECK_B(p, B);
    if (n_batches == 0)
      return HX_OK;
    if (!p->copy_in)
      {
        HX_CUDA(cudaStreamCreateWithFlags(&p->copy_in, cudaStreamNonBlocking));
        HX_CUDA(cudaStreamCreateWithFlags(&p->copy_out, cudaStreamNonBlocking));
        for (int i = 0; i < 6; ++i)
          HX_CUDA(cudaEventCreateWithFlags(&p->pipe_ev[i], cudaEventDisableTiming));
      }
    double *dX[2], *dY[2];
    HX_TRY(p->get_scratch(4, &dX[0], B));
    HX_TRY(p->get_scratch(5, &dY[0], B));
    HX_TRY(p->get_scratch(7, &dX[1], B));
    HX_TRY(p->get_scratch(8, &dY[1], B));
    cudaEvent_t *in_done = p->pipe_ev, *comp_done = p->pipe_ev + 2, *out_done = p->pipe_ev + 4;
    const size_t bytes   = (size_t)p->n_local * B * sizeof(double);
    // nothing of an earlier call may still be using the buffers
    HX_CUDA(cudaEventRecord(comp_done[0], p->stream));
    HX_CUDA(cudaStreamWaitEvent(p->copy_in, comp_done[0], 0));
    for (uint32_t k = 0; k < n_batches; ++k)
      {
        const int s = (int)(k & 1u);
        HX_CHECK(Xh[k] && Yh[k], HX_ERR_INVALID, "null batch pointer");
        if (k >= 2)
          HX_CUDA(cudaStreamWaitEvent(p->copy_in, comp_done[s], 0)); // the filter of batch k-2 is done with this X buffer
        HX_CUDA(cudaMemcpyAsync(dX[s], Xh[k], bytes, cudaMemcpyHostToDevice, p->copy_in));
        HX_CUDA(cudaEventRecord(in_done[s], p->copy_in));
        HX_CUDA(cudaStreamWaitEvent(p->stream, in_done[s], 0));
        if (k >= 2)
          HX_CUDA(cudaStreamWaitEvent(p->stream, out_done[s], 0)); // batch k-2 has left this Y buffer
        HX_TRY(hx_chebyshev_filter(A, BInv, dX[s], dY[s], B, degree, a0, a, b));
        HX_CUDA(cudaEventRecord(comp_done[s], p->stream));
        HX_CUDA(cudaStreamWaitEvent(p->copy_out, comp_done[s], 0));
        HX_CUDA(cudaMemcpyAsync(Yh[k], dY[s], bytes, cudaMemcpyDeviceToHost, p->copy_out));
        HX_CUDA(cudaEventRecord(out_done[s], p->copy_out));
      }
    HX_CUDA(cudaStreamSynchronize(p->copy_out));
    HX_TRY(plan_sync(p));
    return HX_OK;
  }

  int
  hx_residual_chebyshev_filter(hx_op *A, hx_op *Bop, hx_op *BInv, const double *eigenvalues, double *X, double *Y,
                               uint32_t B, uint32_t degree, double a0, double a, double b)
  {
    HX_CHECK(A && Bop && BInv && eigenvalues && X && Y, HX_ERR_INVALID, "null argument");
    hx_plan *p = A->plan;
    HX_CHECK(p == Bop->plan && p == BInv->plan, HX_ERR_INVALID, "operators belong to different plans");
    HX_CHECK(X != Y, HX_ERR_INVALID, "X and Y must not alias (the recurrence ping-pongs between them)");
    HX_CHECK_B(p, B);
    const double e      = 0.5 * (b - a);
    const double c      = 0.5 * (b + a);
    double       sigma  = e / (a0 - c);
    const double sigma1 = sigma;
    const double gamma  = 2.0 / sigma1;
    const size_t nown   = (size_t)p->n_owned * B;
    const size_t nloc   = (size_t)p->n_local * B;
    double *     s1, *s2, *s3, *Res, *ResNew;
    HX_TRY(p->get_scratch(0, &s1));
    HX_TRY(p->get_scratch(1, &s2));
    HX_TRY(p->get_scratch(2, &s3));
    HX_TRY(p->get_scratch(3, &Res));
    HX_TRY(p->get_scratch(6, &ResNew));
    // per-column scalars on the host (the reference keeps them in MemoryStorage and filters them with the
    // same blas kernels, ChebyshevFilter.t.cpp:286-326,400-423); device copies: ones | ev | ev2
    std::vector<double> ones(B, 1.0), ev(eigenvalues, eigenvalues + B), ev1(B, 1.0), ev2(ev);
    HX_TRY(p->ensure_small(3 * (size_t)B));
    double *d_ones = p->d_small.p, *d_ev = p->d_small.p + B, *d_ev2 = p->d_small.p + 2 * (size_t)B;
    HX_CUDA(cudaMemcpyAsync(d_ones, ones.data(), B * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    HX_CUDA(cudaMemcpyAsync(d_ev, ev.data(), B * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    HX_TRY(plan_sync(p));
    double alpha1 = sigma1 / e, alpha2 = -c;
    HX_TRY(op_apply(Bop, X, Y, B, 1, 0));
    HX_TRY(op_apply(A, X, s3, B, 1, 0));
    HX_TRY(launch_axpby_blocked(p, p->n_owned, B, 1.0, d_ones, s3, -1.0, d_ev, Y, Y)); // Y = AX - lambda BX
    HX_CUDA(cudaMemcpyAsync(ResNew, Y, nloc * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    HX_CUDA(cudaMemsetAsync(Res, 0, nloc * sizeof(double), p->stream));
    for (uint32_t j = 0; j < B; ++j)
      ev2[j] = 1.0 * 1.0 * (alpha1 * alpha2) + alpha1 * ev[j] * ev1[j];
    HX_TRY(launch_axpby(p, nown, alpha1, ResNew, 0.0, ResNew, ResNew)); // ascale
    for (uint32_t deg = 2; deg <= degree; ++deg)
      {
        const double sigma2 = 1.0 / (gamma - sigma);
        alpha1              = 2.0 * sigma2 / e;
        alpha2              = -(sigma * sigma2);
        HX_TRY(op_apply(BInv, ResNew, s1, B, 1, 0));
        HX_TRY(op_apply(A, s1, s2, B, 0, 0));
        HX_TRY(launch_axpby(p, nown, alpha1, s2, -c * alpha1, ResNew, s1));
        HX_TRY(launch_axpby(p, nown, 1.0, s1, alpha2, Res, Res));
        HX_CUDA(cudaMemcpyAsync(d_ev2, ev2.data(), B * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        HX_TRY(launch_axpby_blocked(p, p->n_owned, B, 1.0, d_ones, Res, alpha1, d_ev2, Y, Res));
        HX_TRY(plan_sync(p)); // ev2 (host) is rewritten below
        for (uint32_t j = 0; j < B; ++j)
          {
            ev1[j] = (-c * alpha1) * ev2[j] + alpha2 * ev1[j];
            ev1[j] = 1.0 * 1.0 * ev1[j] + alpha1 * ev[j] * ev2[j];
          }
        std::swap(ResNew, Res);
        std::swap(ev1, ev2);
        sigma = sigma2;
      }
    HX_TRY(op_apply(BInv, ResNew, Res, B, 1, 1));
    HX_CUDA(cudaMemcpyAsync(d_ev2, ev2.data(), B * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    HX_TRY(launch_axpby_blocked(p, p->n_owned, B, 1.0, d_ones, Res, 1.0, d_ev2, X, Y));
    HX_TRY(plan_sync(p));
    return HX_OK;
  }

  // --------------------------------------------------------------------- subspace projections ----
  int
  hx_xtopx(hx_op *op, double *X, uint32_t B, uint32_t batch, double *S_host)
  {
    HX_CHECK(op && X && S_host, HX_ERR_INVALID, "null argument");
    hx_plan *p = op->plan;
    HX_CHECK_B(p, B);
    HX_CHECK(batch >= 1, HX_ERR_INVALID, "batch must be >= 1");
    batch = std::min(batch, B);
    double *xin, *xout;
    HX_TRY(p->get_scratch(2, &xin));
    HX_TRY(p->get_scratch(3, &xout));
    HX_TRY(p->ensure_small(gram_workspace_doubles(p, B, batch, p->n_owned))); // S block + split-K partials (gram_block)
    HX_TRY(p->ensure_pinned((size_t)B * batch * sizeof(double)));
    for (size_t i = 0; i < (size_t)B * B; ++i)
      S_host[i] = 0.0;
    for (uint32_t j0 = 0; j0 < B; j0 += batch)
      {
        const uint32_t b = std::min(batch, B - j0);
        double *xb = (b == B) ? X : xin; // whole block: applied in place (copy out / copy back is the identity)
        if (xb != X)
          HX_TRY(copy_cols(p, X, B, j0, xin, b, 0, b, p->n_local));
        HX_TRY(op_apply(op, xb, xout, b, 1, 0));
        double *Sd = p->d_small.p;
        HX_TRY(gram_block(p, X, B, j0, xout, b, p->n_owned, Sd));
        if (p->nranks > 1)
          HX_TRY(comm_allreduce_sum(p->comm, p->stream, Sd, (size_t)(B - j0) * b));
        HX_CUDA(cudaMemcpyAsync(p->h_pinned, Sd, (size_t)(B - j0) * b * sizeof(double), cudaMemcpyDeviceToHost,
                                p->stream));
        // the reference copies the (possibly constraint-filled) batch back into X
        if (xb != X)
          HX_TRY(copy_cols(p, xin, b, 0, X, B, j0, b, p->n_local));
        HX_TRY(plan_sync(p));
        for (uint32_t i = 0; i < b; ++i)
          for (uint32_t j = j0 + i; j < B; ++j)
            S_host[(size_t)j + (size_t)(i + j0) * B] = p->h_pinned[(size_t)i * (B - j0) + (j - j0)];
      }
    return HX_OK;
  }

  int
  hx_subspace_rotation(hx_plan *plan, double *X, uint32_t B, const double *Q_host, int transpose, int lowerTri)
  {
    HX_CHECK(plan && X && Q_host, HX_ERR_INVALID, "null argument");
    HX_CHECK_B(plan, B);
    // effective K x N operand, row-major: Qeff[i*B + j] = Q(i,j) (transpose) or Q(j,i); lower-triangular
    // rotation matrices only couple i <= j-block end, the zeros are kept explicit
    std::vector<double> qeff((size_t)B * B);
    for (uint32_t i = 0; i < B; ++i)
      for (uint32_t j = 0; j < B; ++j)
        qeff[(size_t)i * B + j] = transpose ? Q_host[(size_t)i + (size_t)j * B] : Q_host[(size_t)j + (size_t)i * B];
    HX_TRY(plan->ensure_small((size_t)B * B));
    HX_CUDA(cudaMemcpyAsync(plan->d_small.p, qeff.data(), qeff.size() * sizeof(double), cudaMemcpyHostToDevice,
                            plan->stream));
    HX_TRY(plan_sync(plan));
    double *tmp;
    HX_TRY(plan->get_scratch(2, &tmp));
    return rotate(plan, X, B, plan->n_owned, plan->d_small.p, transpose, lowerTri, tmp);
  }

  int
  hx_l2_norms(hx_plan *plan, const double *X, uint32_t B, double *norms_host)
  {
    HX_CHECK(plan && X && norms_host, HX_ERR_INVALID, "null argument");
    HX_CHECK_B(plan, B);
    HX_CHECK(B <= 256, HX_ERR_UNSUPPORTED, "hx_l2_norms supports B <= 256 per call (batch the columns)");
    HX_TRY(plan->ensure_small((size_t)600 * B + 2 * B));
    double *out = plan->d_small.p + (size_t)600 * B;
    HX_TRY(launch_colsumsq(plan, X, B, plan->n_owned, out));
    if (plan->nranks > 1)
      HX_TRY(comm_allreduce_sum(plan->comm, plan->stream, out, B));
    HX_CUDA(cudaMemcpyAsync(norms_host, out, B * sizeof(double), cudaMemcpyDeviceToHost, plan->stream));
    HX_TRY(plan_sync(plan));
    for (uint32_t j = 0; j < B; ++j)
      norms_host[j] = sqrt(norms_host[j]);
    return HX_OK;
  }

  int
  hx_axpby(hx_plan *plan, uint32_t n_rows, uint32_t B, double alpha, const double *x, double beta, const double *y,
           double *z)
  {
    HX_CHECK(plan && x && y && z, HX_ERR_INVALID, "null argument");
    return launch_axpby(plan, (size_t)n_rows * B, alpha, x, beta, y, z);
  }
  int
  hx_axpby_blocked(hx_plan *plan, uint32_t n_rows, uint32_t B, double alpha1, const double *alpha_host,
                   const double *x, double beta1, const double *beta_host, const double *y, double *z)
  {
    HX_CHECK(plan && x && y && z && alpha_host && beta_host, HX_ERR_INVALID, "null argument");
    HX_TRY(plan->ensure_small(2 * (size_t)B));
    HX_CUDA(cudaMemcpyAsync(plan->d_small.p, alpha_host, B * sizeof(double), cudaMemcpyHostToDevice, plan->stream));
    HX_CUDA(cudaMemcpyAsync(plan->d_small.p + B, beta_host, B * sizeof(double), cudaMemcpyHostToDevice, plan->stream));
    HX_TRY(plan_sync(plan));
    return launch_axpby_blocked(plan, n_rows, B, alpha1, plan->d_small.p, x, beta1, plan->d_small.p + B, y, z);
  }

  int
  hx_plan_trace(hx_plan *plan, int on)
  {
    HX_CHECK(plan, HX_ERR_INVALID, "null plan");
    HX_TRY(plan_sync(plan));
    for (auto &m : plan->trace_marks)
      plan->trace_pool.push_back(m.second);
    plan->trace_marks.clear();
    plan->trace = on != 0;
    return HX_OK;
  }

  int
  hx_plan_trace_report(hx_plan *plan, char *buf, size_t buf_bytes)
  {
    HX_CHECK(plan && buf && buf_bytes > 0, HX_ERR_INVALID, "null argument");
    HX_TRY(plan_sync(plan));
    std::vector<std::pair<std::string, std::pair<double, uint64_t>>> acc; // phase -> (ms, count), first-seen order
    for (size_t i = 1; i < plan->trace_marks.size(); ++i)
      {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, plan->trace_marks[i - 1].second, plan->trace_marks[i].second) != cudaSuccess)
          continue;
        const std::string name = plan->trace_marks[i].first;
        if (name.size() > 6 && name.compare(name.size() - 6, 6, ":begin") == 0)
          continue; // gap before a phase begins belongs to the caller
        auto it = std::find_if(acc.begin(), acc.end(), [&](auto &kv) { return kv.first == name; });
        if (it == acc.end())
          acc.push_back({name, {ms, 1}});
        else
          it->second.first += ms, it->second.second++;
      }
    std::string out = "{";
    for (size_t i = 0; i < acc.size(); ++i)
      {
        char tmp[160];
        snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"ms\": %.6f, \"n\": %llu}", i ? ", " : "", acc[i].first.c_str(), acc[i].second.first,
                 (unsigned long long)acc[i].second.second);
        out += tmp;
      }
    out += "}";
    snprintf(buf, buf_bytes, "%s", out.c_str());
    for (auto &m : plan->trace_marks)
      plan->trace_pool.push_back(m.second);
    plan->trace_marks.clear();
    return HX_OK;
  }

  int
  hx_plan_launch_count(hx_plan *plan, uint64_t *n)
  {
    HX_CHECK(plan && n, HX_ERR_INVALID, "null argument");
    *n = plan->launches;
    return HX_OK;
  }
  int
  hx_plan_enable_kernel_timing(hx_plan *plan, int on)
  {
    HX_CHECK(plan, HX_ERR_INVALID, "null plan");
    plan->timing        = on != 0;
    plan->ev_used       = 0;
    plan->cell_launches = 0;
    return HX_OK;
  }
  int
  hx_plan_cell_kernel_time_ms(hx_plan *plan, double *ms, uint64_t *launches)
  {
    HX_CHECK(plan && ms && launches, HX_ERR_INVALID, "null argument");
    HX_TRY(plan_sync(plan));
    double tot = 0.0;
    for (size_t i = 0; i + 1 < plan->ev_used; i += 2)
      {
        float t = 0.f;
        HX_CUDA(cudaEventElapsedTime(&t, plan->ev_pool[i], plan->ev_pool[i + 1]));
        tot += t;
      }
    plan->ev_used       = 0;
    *ms                 = tot;
    *launches           = plan->cell_launches;
    plan->cell_launches = 0;
    return HX_OK;
  }

  int
  hx_plan_cell_kernel_sm_clock_mhz(hx_plan *plan, double *mhz)
  {
    HX_CHECK(plan && mhz, HX_ERR_INVALID, "null argument");
    HX_TRY(plan_sync(plan));
    unsigned long long c[2] = {0, 0};
    HX_CUDA(cudaMemcpy(c, plan->d_clk.p, sizeof(c), cudaMemcpyDeviceToHost));
    HX_CUDA(cudaMemset(plan->d_clk.p, 0, sizeof(c)));
    *mhz = c[1] ? 1e3 * (double)c[0] / (double)c[1] : 0.0;
    return HX_OK;
  }
}
