"""Headless frames and video: the frame kernels against a host restatement of the same scene
(reference main.cpp:366-466, render.cpp:53-125, shaders.cpp:40-86, postprocess.cu:32-55), the AVI container byte by byte."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

import particlerobotsimulations_b200 as prs
from tests import util


# ---- the container: host code only ---------------------------------------------------------------------------------
def _chunks(buf, pos, end):
    while pos + 8 <= end:
        tag, size = buf[pos:pos + 4], struct.unpack_from("<I", buf, pos + 4)[0]
        yield tag, pos + 8, size
        pos += 8 + size + (size & 1)


def test_video_container_layout(tmp_path):
    """An AVI a parser can walk: RIFF sizes add up, header counts are patched on close, frames come back top-down."""
    w, h, n = 37, 11, 5                     # a width whose rows need padding to 4 bytes
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    path = str(tmp_path / "t.avi")
    vw = prs.VideoWriter(path, w, h, 20.0)
    for f in frames:
        assert vw.write(f) == 0
    assert vw.close() == n
    buf = open(path, "rb").read()
    assert buf[:4] == b"RIFF" and buf[8:12] == b"AVI " and struct.unpack_from("<I", buf, 4)[0] == len(buf) - 8
    top = {tag + buf[p:p + 4] if tag == b"LIST" else tag: (p, s) for tag, p, s in _chunks(buf, 12, len(buf))}
    assert set(top) == {b"LISThdrl", b"LISTmovi", b"idx1"}
    hp, hs = top[b"LISThdrl"]
    hdr = {tag + buf[p:p + 4] if tag == b"LIST" else tag: (p, s) for tag, p, s in _chunks(buf, hp + 4, hp + hs)}
    ap, asz = hdr[b"avih"]
    usec, _, _, flags, total, _, streams, _, aw, ah = struct.unpack_from("<10I", buf, ap)
    assert (asz, usec, flags & 0x10, total, streams, aw, ah) == (56, 50000, 0x10, n, 1, w, h)
    sp, ss = hdr[b"LISTstrl"]
    strl = {tag: (p, s) for tag, p, s in _chunks(buf, sp + 4, sp + ss)}
    p, s = strl[b"strh"]
    assert s == 56 and buf[p:p + 8] == b"vidsDIB "
    scale, rate, _, length = struct.unpack_from("<4I", buf, p + 20)
    assert rate / scale == 20.0 and length == n
    p, s = strl[b"strf"]
    bi_size, bw, bh, planes, bits, comp, image = struct.unpack_from("<IiiHHII", buf, p)
    row = (3 * w + 3) & ~3
    assert (s, bi_size, bw, bh, planes, bits, comp, image) == (40, 40, w, h, 1, 24, 0, row * h)
    mp, ms = top[b"LISTmovi"]
    got = []
    for tag, p, s in _chunks(buf, mp + 4, mp + ms):
        assert tag == b"00db" and s == row * h
        img = np.frombuffer(buf, np.uint8, s, p).reshape(h, row)[:, :3 * w].reshape(h, w, 3)
        got.append(img[::-1])               # bottom-up in the file
    assert np.array_equal(np.stack(got), frames)
    ip, isz = top[b"idx1"]
    assert isz == 16 * n
    for k in range(n):
        tag, fl, off, size = struct.unpack_from("<4sIII", buf, ip + 16 * k)
        assert tag == b"00db" and fl == 0x10 and size == row * h and buf[mp + off:mp + off + 4] == b"00db"


def test_video_open_refuses_nonsense(tmp_path):
    L = prs.lib()
    assert not L.prs_video_open(os.fsencode(str(tmp_path / "a.avi")), 0, 10, 20.0)
    assert not L.prs_video_open(os.fsencode(str(tmp_path / "no_such_dir" / "a.avi")), 8, 8, 20.0)
    assert L.prs_video_write(None, None) == -1 and L.prs_video_close(None) == -1


def test_view_from_camera_is_the_reference_projection():
    """gluPerspective(60) from camera_y straight down: the image height spans 2 camera_y tan(30 deg) of floor"""
    v = prs.view_from_camera(1920, 1080, 10.0, 0.25)
    assert (v.width, v.height, v.center_x, v.center_y) == (1920, 1080, 0.0, 0.0)
    assert abs(v.world_per_pixel * 1080 - 2 * 10.0 * np.tan(np.pi / 6)) < 1e-5 and v.light_radius == 0.25


# ---- the frame: a host restatement of the scene, pixel by pixel -------------------------------------------------------
def frame_restated(view, p, world_half, pos, rad, col):
    """what k_frame_splat + k_frame_resolve compute, with numpy float32 operations in the same order"""
    f = np.float32
    W, H, s = int(view.width), int(view.height), f(view.world_per_pixel)
    wx = ((np.arange(W, dtype=f) + f(0.5)) - f(0.5) * f(W)) * s + f(view.center_x)
    wy = (f(0.5) * f(H) - (np.arange(H, dtype=f) + f(0.5))) * s + f(view.center_y)
    X, Y = np.meshgrid(wx, wy)

    def disc(cx, cy, r):
        dx, dy = X - f(cx), Y - f(cy)
        return (dx * dx + dy * dy) < f(r) * f(r)

    rgb = np.full((H, W, 3), 0.25, f)
    rgb[(np.abs(X) <= f(world_half)) & (np.abs(Y) <= f(world_half))] = 1.0
    key = np.full((2, H, W), -1, np.int64)
    for i in range(len(pos) - 1, -1, -1):     # descending: the lowest index is written last and wins
        x, y, plane = f(pos[i, 0]), f(pos[i, 1]), 0
        if y > f(1000.0):
            y, plane = y - f(2000.0), 1
        if not rad[i] > 0:
            continue
        key[plane][disc(x, y, rad[i])] = i
    hit = key[0] >= 0
    rgb[hit] = col[key[0][hit], :3]
    rgb[disc(p.light_x, p.light_y, view.light_radius)] = (0.8, 0.8, 0.0)
    obs = np.zeros((H, W), bool)
    for i in range(p.n_cir_obstacles):
        obs |= disc(p.x_cir_obs[i], p.y_cir_obs[i], p.r_cir_obs[i])
    for i in range(p.nobstacles):
        x1, x2, y1, y2 = p.x1obs[i], p.x2obs[i], p.y1obs[i], p.y2obs[i]
        obs |= (X >= f(min(x1, x2))) & (X <= f(max(x1, x2))) & (Y >= f(min(y1, y2))) & (Y <= f(max(y1, y2)))
    rgb[obs] = 0.2
    hit = key[1] >= 0
    rgb[hit] = col[key[1][hit], :3]
    b = np.rint(np.clip(rgb, 0, 1) * f(255.0)).astype(np.uint8)
    return b[:, :, ::-1]


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["example.cfg", "example_dead_cells.cfg", "example_obstacle.cfg", "example_gap.cfg"])
def test_frame_equals_host_restatement(cfg):
    """renderFrame after 30 steps == the restated scene, byte for byte (robots coloured by updateCol, dead robots black,
    overlapping sprites resolved to the lowest index, obstacles and the light marker over the robots, the trail on top)"""
    p, o = prs.load_cfg(os.path.join(util.ROOT, "examples", cfg))
    L = prs.lib()
    L.cudaInit(0, None)
    sim = prs.Simulation(p, 64.0, prs.BACKEND_FUSED)
    sim.srand(p.seed)
    sim.reset()
    for _ in range(30):
        sim.update(o.timestep, o.timestep)
    n, trail = int(p.nCells), int(p.centroid_steps)
    view = prs.view_from_camera(320, 180, o.camera_y, o.light_radius)
    got = sim.render_frame(view)
    # the buffers the frame was drawn from: positions and radii with the trail, colours
    pos = np.empty((n + trail, 2), np.float32)
    rad = np.empty(n + trail, np.float32)
    L.copyArrayFromDevice(pos.ctypes.data, sim.device_ptr(prs.POSITION), None, pos.nbytes)
    L.copyArrayFromDevice(rad.ctypes.data, sim.device_ptr(prs.RADII), None, rad.nbytes)
    col = np.ones((n + trail, 4), np.float32)
    col[n:, 1:3] = 0.0
    dcol, dead = C.c_void_p(), sim.get(prs.DEAD)
    L.allocateArray(C.byref(dcol), col.nbytes)
    L.copyArrayToDevice(dcol, col.ctypes.data, 0, col.nbytes)
    L.updateCol(sim.device_ptr(prs.RADII), dcol, n, sim.device_ptr(prs.POSITION), sim.device_ptr(prs.PHASE), sim.device_ptr(prs.DEAD))
    L.copyArrayFromDevice(col.ctypes.data, dcol, None, col.nbytes)
    L.freeArray(dcol)
    assert np.all(col[:n][dead != 0, :3] == 0.0)
    want = frame_restated(view, p, 64.0, pos, rad, col)
    assert got.shape == want.shape == (180, 320, 3)
    assert np.array_equal(got, want), int((got != want).any(2).sum())
    # the scene is there: robots (neither floor white nor background grey nor obstacle grey), and the yellow light marker
    flat = got.reshape(-1, 3)
    assert (flat == (0, 204, 204)).all(1).any()
    robots = ~((flat == 255).all(1) | (flat == 64).all(1) | (flat == 51).all(1) | (flat == (0, 204, 204)).all(1))
    assert robots.sum() > 20
    sim.close()


@pytest.mark.gpu
def test_frame_trail_and_zoomed_out_swarm():
    """trail entries (y + 2000, calcCOG1) land on top of the robots; a 65 536-robot block seen from far away (sub-pixel
    sprites) equals the restatement too"""
    p, o = prs.load_cfg(os.path.join(util.ROOT, "examples", "example.cfg"))
    p.nCells = 256 * 256
    p.centroid_steps = 4
    L = prs.lib()
    L.cudaInit(0, None)
    sim = prs.Simulation(p, 64.0, prs.BACKEND_FUSED)
    sim.init_hex(256, 256, 0.17, 0.01 * p.max_radius, 5555)
    for _ in range(3):
        sim.update(o.timestep, o.timestep)
    n = int(p.nCells)
    # a trail of four marks across the block
    tr = np.array([[-3.0, 2000.0], [0.0, 2001.0], [3.0, 1998.5], [-5000.0, 0.0]], np.float32)
    L.copyArrayToDevice(sim.device_ptr(prs.POSITION), tr.ctypes.data, n * 8, tr.nbytes)
    big = np.full(4, 0.6, np.float32)           # marks of a few pixels at this zoom (the default centroid_radius is sub-pixel here)
    L.copyArrayToDevice(sim.device_ptr(prs.RADII), big.ctypes.data, n * 4, big.nbytes)
    view = prs.view_from_camera(256, 144, 40.0, 1.0)
    got = sim.render_frame(view)
    pos = np.empty((n + 4, 2), np.float32)
    rad = np.empty(n + 4, np.float32)
    L.copyArrayFromDevice(pos.ctypes.data, sim.device_ptr(prs.POSITION), None, pos.nbytes)
    L.copyArrayFromDevice(rad.ctypes.data, sim.device_ptr(prs.RADII), None, rad.nbytes)
    col = np.ones((n + 4, 4), np.float32)
    col[n:, 1:3] = 0.0
    # updateCol's ramp (kernel_impl.cuh:413-415) is checked against the reference's kernel in test_parity_gpu.py
    dcol = C.c_void_p()
    L.allocateArray(C.byref(dcol), col.nbytes)
    L.copyArrayToDevice(dcol, col.ctypes.data, 0, col.nbytes)
    L.updateCol(sim.device_ptr(prs.RADII), dcol, n, sim.device_ptr(prs.POSITION), sim.device_ptr(prs.PHASE), sim.device_ptr(prs.DEAD))
    L.copyArrayFromDevice(col.ctypes.data, dcol, None, col.nbytes)
    L.freeArray(dcol)
    want = frame_restated(view, p, 64.0, pos, rad, col)
    assert np.array_equal(got, want), int((got != want).any(2).sum())
    assert (got.reshape(-1, 3) == (0, 0, 255)).all(1).sum() >= 3       # red trail marks, B, G, R
    sim.close()


@pytest.mark.gpu
def test_runner_writes_video(tmp_path):
    """ParticleBot --video: a frame every DISPLAY_INTERVAL steps, every VIDEO_INTERVAL-th of them in the file"""
    cfg = tmp_path / "v.cfg"
    src = open(os.path.join(util.ROOT, "examples", "example.cfg")).read()
    cfg.write_text(src + "\nDISPLAY_INTERVAL\n10\nVIDEO_INTERVAL\n2\n")
    out = tmp_path / "o.avi"
    exe = os.path.join(util.ROOT, "particlerobotsimulations_b200", "ParticleBot")
    r = subprocess.run([exe, str(cfg), "--steps", "100", "--no-csv", "--quiet", "--video", str(out), "--video-size", "160x90",
                        "--frame-ppm", str(tmp_path / "last.ppm")], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    assert "5 video frames (160x90)" in r.stderr, r.stderr       # steps 0, 20, 40, 60, 80
    buf = out.read_bytes()
    assert buf[:4] == b"RIFF" and struct.unpack_from("<I", buf, 48)[0] == 5
    ppm = (tmp_path / "last.ppm").read_bytes()
    assert ppm.startswith(b"P6\n160 90\n255\n") and len(ppm) == 14 + 160 * 90 * 3
