"""Problem definitions shared by the tests, smoke() and bench.py."""
import numpy as np

import flexpart_b200 as fb


def config_c1(npart=10000, **over):
    """C1: the shipped options/ forward run (options/COMMAND, RELEASES, OUTGRID,
    AGECLASSES; species 24 AIRTRACER) on the synthetic 1 deg x 138-level field.
    LCONVECTION is off because convmix is outside the path."""
    kw = dict(nx=361, ny=181, nz=138, dx=1.0, dy=1.0, xlon0=-180.0, ylat0=-90.0,
              ldirect=1, lsynctime=900, ctl=-5.0, ifine=4,
              outlon0=-25.0, outlat0=10.0, numxgrid=85, numygrid=65, dxout=1.0, dyout=1.0,
              outheights=(100.0, 500.0, 1000.0, 50000.0), lage=(1728000,),
              ioutputforeachrelease=1, npart=(npart,), xmass=[[1.0]], nspec=1)
    kw.update(over)
    return fb.make_config(**kw)


def releases_c1(cb, start=0):
    """options/RELEASES: point release lon 0, lat 20, 50 m agl, instantaneous."""
    return fb.Releases(cb, lon1=0.0, lon2=0.0, lat1=20.0, lat2=20.0, z1=50.0, z2=50.0,
                       start=start, end=start)


def config_small(nrel=4, npart_each=500, nx=73, ny=37, nz=40, **over):
    """Coarse global grid (5 deg x 40 levels) that the oracle steps in seconds."""
    kw = dict(nx=nx, ny=ny, nz=nz, dx=360.0 / (nx - 1), dy=180.0 / (ny - 1), xlon0=-180.0, ylat0=-90.0,
              ldirect=1, lsynctime=900, ctl=5.0, ifine=4,
              outlon0=-180.0, outlat0=-90.0, numxgrid=72, numygrid=36, dxout=5.0, dyout=5.0,
              outheights=(100.0, 1000.0, 5000.0, 50000.0), lage=(1728000,),
              ioutputforeachrelease=0, npart=(npart_each,) * nrel, nspec=1,
              height=fb.synth_heights(138)[::3][:nz] if nz <= 46 else None)
    kw.update(over)
    return fb.make_config(**kw)


def releases_boxes(cb, seed=1, zmax=2000.0, lat_range=(-60.0, 60.0), start=0, end=0, width=5.0):
    """numpoint box releases at seeded positions, 0..zmax m agl."""
    n = cb.cfg.numpoint
    r = np.random.RandomState(seed)
    lon = r.uniform(-170.0, 170.0 - width, n)
    lat = r.uniform(lat_range[0], lat_range[1] - width, n)
    return fb.Releases(cb, lon1=lon, lon2=lon + width, lat1=lat, lat2=lat + width,
                       z1=np.zeros(n), z2=np.full(n, zmax), start=np.full(n, start), end=np.full(n, end))


def seeded_particles(cb, n, seed=7, zmax=3000.0, lat_range=(-80.0, 80.0), itime=0, nspec=None):
    """n particles at seeded positions, all active at `itime` and new (itramem = itime)."""
    c = cb.cfg
    nspec = nspec or c.nspec
    p = fb.Particles(c.maxpart, nspec)
    r = np.random.RandomState(seed)
    p.xtra1[:n] = r.uniform(0.0, c.nxmin1 - 1e-3, n)
    y0 = (lat_range[0] - c.ylat0) / c.dy
    y1 = (lat_range[1] - c.ylat0) / c.dy
    p.ytra1[:n] = r.uniform(y0, y1, n)
    p.ztra1[:n] = r.uniform(1.0, zmax, n).astype(np.float32)
    p.itra1[:n] = itime
    p.itramem[:n] = itime
    p.npoint[:n] = r.randint(1, c.numpoint + 1, n)
    p.nclass[:n] = 1
    p.idt[:n] = c.mintime
    p.xmass1[:n, :] = 1.0
    p.numpart = n
    return p


def met_pair(cb, t0=0, t1=10800):
    return fb.MetFields(cb).synth(t0), fb.MetFields(cb).synth(t1)
