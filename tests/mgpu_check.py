"""Multi-GPU check, run as  torchrun --nproc-per-node N tests/mgpu_check.py  (one rank per GPU, NCCL).

Every rank steps its z-slab of a scene whose particles cross the slab boundaries, and ALSO steps the whole scene on
its own GPU with a single-rank context; after each step the rank's own particles must be exactly the whole-scene
particles that lie in its slab (same count; positions / velocities within the solver tolerance), and the rank's
cells must match the whole-grid cells.  Exit code 0 = pass."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from scipy.spatial import cKDTree
    from libfluid_b200 import capi, slabs

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("MGPU_HANG_DUMP_S", "75")), exit=True)  # a hang names its line
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    failures = []
    # two layouts: slabs of unequal thickness and an odd depth (no slab-aligned coarsening: level 0 is the only
    # distributed multigrid level), and even slabs, where the distributed hierarchy -- and, with LFK_TUNE=mg_agg=1, the
    # agglomerated coarse levels -- are exercised.  MGPU_ALIGNED=0 / 1 restricts the run to one of them.
    only = os.environ.get("MGPU_ALIGNED")
    layouts = [(capi.APIC, False), (capi.FLIP, False), (capi.APIC, True)]
    if only in ("0", "1"):
        layouts = [(capi.APIC, only == "1"), (capi.FLIP, only == "1")]
    for method, aligned in layouts:
        box = [capi.nccl_unique_id() if rank == 0 else None]  # one NCCL id per communicator
        dist.broadcast_object_list(box, src=0)
        n = (24, 20, 16 * world) if aligned else (24, 20, 8 * world + 3)
        kw = dict(cell_size=1.0, gravity=(0.0, -981.0, 0.0), method=method, blending_factor=0.95, max_iterations=2000)
        multi = capi.Context(n, device=local, nranks=world, rank=rank, nccl_id=box[0], **kw)
        whole = capi.Context(n, device=local, **kw)
        z0, z1 = multi.slab()
        assert (z0, z1 - z0) == slabs.slab_range(n[2], world, rank)
        # a slanted block of water moving along +z / -z so that particles migrate both ways every step
        for ctx in (multi, whole):
            ctx.seed_box_device((2.0, 1.0, 1.0), (14.0, 9.0, n[2] * 0.45), velocity=(3.0, 0.0, 55.0), density=2, seed=7)
            ctx.seed_box_device((6.0, 3.0, n[2] * 0.55), (15.0, 12.0, n[2] * 0.4), velocity=(-2.0, 0.0, -48.0),
                                density=2, seed=9, append=True)
        # FLIP / PIC leave the APIC c rows untouched, so they travel with the particle through sorts and exchanges: the
        # check tags cx[0] with a particle id there and matches particles by id.  (APIC overwrites c every step, but
        # particles that coincide -- clamped into the same corner -- then also agree in every field, so matching by
        # position is unambiguous for everything that is compared.)
        by_id = method != capi.APIC
        if by_id:
            allp = whole.download_particles()
            allp["cx"][:, 0] = np.arange(allp.shape[0], dtype=np.float64)
            allp["cy"] = allp["position"]  # a second carried tag: where the particle started
            whole.upload_particles(allp)
            zc0 = slabs.z_cell(allp["position"][:, 2], n[2])
            multi.upload_particles(allp[(zc0 >= z0) & (zc0 < z1)])
        total = torch.tensor([multi.num_particles()], dtype=torch.int64, device="cuda")
        dist.all_reduce(total)
        if int(total.item()) != whole.num_particles():
            failures.append("seeding: %d vs %d" % (int(total.item()), whole.num_particles()))
        moved = 0
        for step in range(6):
            dt = 0.02
            multi.time_step(dt)
            whole.time_step(dt)
            moved += multi.stats()["exchanged_particles"]
            a, b = multi.download_particles(), whole.download_particles()
            # ownership is settled by the sort (after advection); the position correction may then nudge a boundary
            # particle one cell across, so a rank's particles are matched against the WHOLE scene, and the ranks'
            # sets must partition it
            tag = "method %d%s step %d rank %d" % (method, " aligned" if aligned else "", step, rank)
            cnt = torch.tensor([a.shape[0]], dtype=torch.int64, device="cuda")
            dist.all_reduce(cnt)
            if int(cnt.item()) != b.shape[0]:
                failures.append("%s: ranks hold %d particles, whole-scene run %d" % (tag, int(cnt.item()), b.shape[0]))
                continue
            if a.shape[0]:
                zc = slabs.z_cell(a["position"][:, 2], n[2])
                if zc.min() < z0 - 1 or zc.max() > z1:
                    failures.append("%s: own particle outside slab +- 1 (z cells %d..%d)" % (tag, zc.min(), zc.max()))
                    continue
                vmax = max(1.0, np.abs(b["velocity"]).max())
                if by_id:
                    for nm, q in (("slab", a), ("whole-scene", b)):  # the whole payload must have travelled together
                        qi = np.clip(np.rint(q["cx"][:, 0]).astype(np.int64), 0, allp.shape[0] - 1)
                        torn = (q["cy"] != allp["position"][qi]).any(axis=1)
                        if torn.any():
                            failures.append("%s: %s run: %d particles carry a torn payload" % (tag, nm, int(torn.sum())))
                    ids = np.rint(a["cx"][:, 0]).astype(np.int64)
                    order = np.argsort(np.rint(b["cx"][:, 0]).astype(np.int64))
                    if np.unique(ids).size != ids.size or ids.min() < 0 or ids.max() >= b.shape[0]:
                        failures.append("%s: particle ids are not a subset of the whole scene's" % tag)
                        continue
                    bids = np.rint(b["cx"][:, 0]).astype(np.int64)
                    if not np.array_equal(bids[order], np.arange(b.shape[0])):
                        failures.append("%s: the whole-scene run lost its particle ids" % tag)
                        continue
                    idx = order[ids]
                    d = np.abs(a["position"] - b["position"][idx]).max(axis=1)
                    if d.max() > 1e-6:
                        w = int(np.argmax(d))
                        failures.append("%s: positions differ by id (max %.3e; %d of %d particles off; worst id %d at %s vs %s, "
                                        "z cell %d, slab %d..%d)" % (tag, d.max(), int((d > 1e-6).sum()), d.size, ids[w],
                                                                     a["position"][w], b["position"][idx[w]], zc[w], z0, z1))
                        continue
                else:
                    d, idx = cKDTree(b["position"]).query(a["position"])
                    # (particles clamped into the same wall corner coincide, so the match need not be injective there)
                    if np.unique(idx).size < 0.995 * idx.size or d.max() > 1e-6:
                        failures.append("%s: positions differ (max %.3e, unique %d / %d)" % (tag, d.max(), np.unique(idx).size, idx.size))
                        continue
                dvv = np.abs(a["velocity"] - b["velocity"][idx]).max(axis=1)
                dv = dvv.max()
                if dv > 1e-4 * vmax:
                    w = int(np.argmax(dvv))
                    zb = zc[dvv > 1e-4 * vmax]
                    failures.append("%s: velocities differ (max %.3e at pos %s, v %s vs %s; %d bad, z cells %d..%d; slab %d..%d)"
                                    % (tag, dv, a["position"][w], a["velocity"][w], b["velocity"][idx[w]], zb.size,
                                       zb.min(), zb.max(), z0, z1))
                if not by_id:
                    dc = max(np.abs(a[f] - b[f][idx]).max() for f in ("cx", "cy", "cz"))
                    if dc > 1e-4 * vmax:
                        failures.append("%s: APIC c rows differ (max %.3e)" % (tag, dc))
                        continue
            ca, cb = multi.download_cells(), whole.download_cells()
            own = slice(z0 * n[0] * n[1], z1 * n[0] * n[1])
            if not np.array_equal(ca["type"][own], cb["type"][own]):
                failures.append("%s: cell types differ" % tag)
            dvel = np.abs(ca["vel"][own] - cb["vel"][own]).max(axis=1)
            if dvel.max() > 1e-4 * max(1.0, np.abs(cb["vel"]).max()):
                badc = np.nonzero(dvel > 1e-4 * max(1.0, np.abs(cb["vel"]).max()))[0] + own.start
                zz = badc // (n[0] * n[1])
                failures.append("%s: face velocities differ (max %.3e; %d cells, z %d..%d; slab %d..%d)"
                                % (tag, dvel.max(), badc.size, zz.min(), zz.max(), z0, z1))
        if rank == 0:
            print("method %d%s: %d steps compared, %d particles exchanged by rank 0"
                  % (method, " (aligned slabs)" if aligned else "", step + 1, moved), flush=True)
        ex = torch.tensor([moved], dtype=torch.int64, device="cuda")
        dist.all_reduce(ex)
        if int(ex.item()) == 0:
            failures.append("method %d: no particle was ever exchanged -- the scene does not exercise migration" % method)
        multi.close()
        whole.close()
    bad = torch.tensor([len(failures)], dtype=torch.int64, device="cuda")
    dist.all_reduce(bad)
    for f in failures:
        print("[rank %d] FAIL %s" % (rank, f), flush=True)
    if rank == 0:
        print("mgpu_check: %s (%d ranks)" % ("ok" if int(bad.item()) == 0 else "FAILED", world), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(bad.item()) == 0 else 1)


if __name__ == "__main__":
    main()
