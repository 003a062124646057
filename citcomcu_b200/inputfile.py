"""Reference-format input files (``name=value`` lines, parsed by the reference's
Parsing.c:101 ``setup_parser``) for the workloads of BASELINE.json.

The host keeps the reference's input format (north star), so the same text
drives the oracle (`oracle/_ref/ref_harness`), the drop-in run and our own
host mirror (`citcomcu_b200.problem.CartesianProblem.from_input`).

Key names are the reference's (Instructions.c:690-1120, Viscosity_structures.c:57-326,
Advection_diffusion.c:64-113, Convection.c:51-170, Nodal_mesh.c:98-125).
"""
from __future__ import annotations

_DEFAULTS = dict(
    datafile="out/run", use_scratch="local", oldfile="out/run",
    restart=0, restart_timesteps=0, stokes_flow_only=0, maxstep=5, storage_spacing=1000000,
    Solver="multigrid", node_assemble=1,
    rayleigh=30000, rayleigh_comp=0, composition=0, Q0=0, Q0_enriched=0,
    markers_per_ele=0, comp_depth=0.605, visc_heating=0, adi_heating=0,
    nprocx=1, nprocz=1, nprocy=1,
    nodex=33, nodez=33, nodey=17,
    mgunitx=2, mgunitz=2, mgunity=1, levels=5,
    Geometry="cart3d",
    dimenx=1.0079, dimenz=1.0, dimeny=0.6283,
    z_grid_layers=2, zz="0.0,1.0", nz="1,33",
    x_grid_layers=2, xx="0,1.0079", nx="1,33",
    y_grid_layers=2, yy="0,0.6283", ny="1,17",
    z_lmantle=0.76655052, z_410=0.857143, z_lith=0.9651568,
    rheol=0, TDEPV="off", VISC_UPDATE="off", update_every_steps=1,
    num_mat=4, visc0="1,1,1,1", viscE="0,0,0,0", viscT="273,273,273,273",
    viscZ="5e-6,5e-6,5e-6,5e-6",
    SDEPV="off", sdepv_misfit=0.010, sdepv_expt="1,1,1,1", sdepv_trns="1.e0,1.e0,1.e0,1.e0",
    VMIN="off", visc_min=5.0e-2, VMAX="off", visc_max=2.0e04,
    visc_smooth_cycles=1, Viscosity="system",
    layerd=2870000.0, radius=6370000.0, ReferenceT=3800.0, refvisc=1.0e20,
    density=3300.0, thermdiff=1.0e-6, gravacc=9.8, thermexp=5e-5, cp=1250,
    wdensity=0.0, visc_factor=1.0, thermexp_factor=1.0, thermdiff_factor=1.00,
    dissipation_number=0.0, surf_temp=0.078947,
    Ra_410=0.0, Ra_670=0.0, clapeyron410=3.0e6, clapeyron670=-3.0e6,
    width410=3.5e4, width670=3.5e4,
    topvbc=1, topvbxval=0.0, topvbyval=0.0, botvbc=1, botvbxval=0.0, botvbyval=0.0,
    toptbc=1, bottbc=1, toptbcval=0.0, bottbcval=1.0,
    periodicx="off", periodicy="off", flowthroughx="off", flowthroughy="off",
    num_perturbations=1, perturbmag=0.001, perturbk=1.0, perturbl=6.0, perturbm=0.0,
    Problem="convection",
    aug_lagr="on", aug_number=1.0e3, precond="on", orthogonal="off", maxsub=1,
    viterations=2, mg_cycle=1, down_heavy=3, up_heavy=3, vlowstep=20, vhighstep=3,
    piterations=375, accuracy=1.0e-3, tole_compressibility=1e-7,
    adv_sub_iterations=2, finetunedt=0.75, ll_max=20, nlong=180, nlati=90,
    DESCRIBE="off", BEGINNER="off", VERBOSE="off", verbose="off", COMPRESS="off",
    see_convergence=1,
)


def make_input(**overrides) -> str:
    """Return input-file text; `overrides` replace/add keys."""
    params = dict(_DEFAULTS)
    params.update(overrides)
    return "".join(f"{k}={v}\n" for k, v in params.items())


def cartesian_box(elx: int, ely: int, elz: int, levels: int, *, dimx=1.0, dimy=1.0, dimz=1.0,
                  nproc=(1, 1, 1), **overrides) -> str:
    """Uniform Cartesian box of elx*ely*elz elements (x, y, z) with `levels` multigrid levels.

    mgunit = el / 2**(levels-1) per direction (README:149-151 requires divisibility).
    """
    f = 2 ** (levels - 1)
    for n, name in ((elx, "x"), (ely, "y"), (elz, "z")):
        if n % f:
            raise ValueError(f"el{name}={n} not divisible by 2**(levels-1)={f}")
    p = dict(
        mgunitx=elx // f, mgunity=ely // f, mgunitz=elz // f, levels=levels,
        nodex=elx + 1, nodey=ely + 1, nodez=elz + 1,
        dimenx=dimx, dimeny=dimy, dimenz=dimz,
        xx=f"0,{dimx}", nx=f"1,{elx + 1}", yy=f"0,{dimy}", ny=f"1,{ely + 1}",
        zz=f"0.0,{dimz}", nz=f"1,{elz + 1}",
        nprocx=nproc[0], nprocy=nproc[1], nprocz=nproc[2],
    )
    p.update(overrides)
    return make_input(**p)


def busse1a(levels: int = 5, maxstep: int = 5, **overrides) -> str:
    """BASELINE config 1: Busse et al. 1993 case 1a (isoviscous, no-slip top/bottom).

    levels=5 is the reference's 32x32(z)x16(y) mesh; smaller `levels` give the same
    physics on 2x1x2 * 2**(levels-1) elements for fast tests.
    """
    f = 2 ** (levels - 1)
    return cartesian_box(2 * f, 1 * f, 2 * f, levels, dimx=1.0079, dimy=0.6283, dimz=1.0,
                         maxstep=maxstep, **overrides)


def tdepv_box(elx: int, ely: int, elz: int, levels: int, *, nproc=(1, 1, 1), **overrides) -> str:
    """BASELINE config 3 recipe: Ra=1e7, viscosity contrast 1e5 (rheol=0,
    viscE=ln 1e5), free slip, aspect ratio 2x2x1, K rebuilt every step."""
    p = dict(
        rayleigh=1e7, rheol=0, TDEPV="on", VISC_UPDATE="on", update_every_steps=1,
        viscE="11.512925,11.512925,11.512925,11.512925", visc0="1,1,1,1",
        VMIN="off", VMAX="off", topvbc=0, botvbc=0, perturbmag=0.01, perturbl=1.0,
        mg_cycle=1, down_heavy=3, up_heavy=3, vlowstep=20, vhighstep=3, piterations=375,
        accuracy=1e-3, adv_sub_iterations=2, finetunedt=0.75,
    )
    p.update(overrides)
    return cartesian_box(elx, ely, elz, levels, dimx=2.0, dimy=2.0, dimz=1.0, nproc=nproc, **p)


def input1_cart(levels: int = 4, maxstep: int = 5, **overrides) -> str:
    """BASELINE config 2: examples/input1 as a single-rank Cartesian run -- unit box, mgunit 6x6x6
    (48^3 elements at levels=4), NON-UNIFORM z spacing (refined boundary layers: zz=0,0.1,0.9,1 / nz=1,7,43,49),
    Ra=10.97394e5, constant viscosity with the cutoffs on, free slip, augmented Lagrangian 1e3 with the preconditioner.
    Smaller `levels` keep the same layer proportions (1/8, 3/4, 1/8 of the elements)."""
    n = 6 * 2 ** (levels - 1)
    a = n // 8
    p = dict(
        rayleigh=10.97394e5, TDEPV="off", VISC_UPDATE="off", update_every_steps=2, viscE="6.9077553,6.9077553,6.9077553,6.9077553",
        VMIN="on", visc_min=5.0e-2, VMAX="on", visc_max=2.0e04, topvbc=0, botvbc=0, perturbmag=0.001, perturbk=1.0, perturbl=6.0,
        aug_lagr="on", aug_number=1.0e3, precond="on", dissipation_number=2.601, accuracy=1.0e-3,
        z_grid_layers=4, zz="0.0,0.1,0.9,1.0", nz=f"1,{1 + a},{1 + n - a},{1 + n}",
    )
    p.update(overrides)
    return cartesian_box(n, n, n, levels, dimx=1.0, dimy=1.0, dimz=1.0, maxstep=maxstep, **p)


def input1_rsphere(levels: int = 3, maxstep: int = 2, **overrides) -> str:
    """BASELINE config 4 geometry: examples/input1's regional-spherical block (radius 0.55 .. 1, colatitude 73.5 .. 106.5 deg,
    longitude 0 .. 36 deg, refined radial boundary layers) as a single-rank run with mgunit 6x6x6 and `levels` levels."""
    n = 6 * 2 ** (levels - 1)
    a = n // 8
    p = dict(
        Geometry="Rsphere", rayleigh=10.97394e5, TDEPV="off", VISC_UPDATE="off", update_every_steps=2, viscE="6.9077553,6.9077553,6.9077553,6.9077553",
        VMIN="on", visc_min=5.0e-2, VMAX="on", visc_max=2.0e04, topvbc=0, botvbc=0, perturbmag=0.001, perturbk=1.0, perturbl=6.0,
        aug_lagr="on", aug_number=1.0e3, precond="on", dissipation_number=2.601, accuracy=1.0e-3,
        radius_inner=0.55, radius_outer=1.0, theta_north=73.5, theta_south=106.5, fi_west=0, fi_east=36.0,
        r_grid_layers=4, rr="0.55,0.59,0.96,1.0", nr=f"1,{1 + a},{1 + n - a},{1 + n}",
        t_grid_layers=2, tt="73.5,106.5", nt=f"1,{1 + n}", f_grid_layers=2, ff="0,36", nf=f"1,{1 + n}",
        r_lmantle=0.89482, r_410=0.9356358, r_lith=0.984301,
        z_grid_layers=4, zz="0.0,0.1,0.9,1.0", nz=f"1,{1 + a},{1 + n - a},{1 + n}",
    )
    p.update(overrides)
    return cartesian_box(n, n, n, levels, dimx=1.0, dimy=1.0, dimz=1.0, maxstep=maxstep, **p)
