"""CPU: the image export of SURVEY 8(f)4 (render.py) -- colormap anchors, geometry and orientation of the JPG that
stands where the reference's matplotlib savefig does (meteor_detect_class/prime_detection.py:96-105)."""
import numpy as np

from meteor_scatter_b200 import render


def test_viridis_polynomial_hits_the_matplotlib_anchor_colours():
    # matplotlib's viridis table at 0, 0.25, 0.5, 0.75, 1 (8-bit)
    ref = np.array([[68, 1, 84], [59, 82, 139], [33, 145, 140], [94, 201, 98], [253, 231, 37]])
    got = render.viridis(np.array([0.0, 0.25, 0.5, 0.75, 1.0])).astype(int)
    assert np.abs(got - ref).max() <= 5
    assert np.array_equal(render.viridis(np.array([-3.0, 7.0])), render.viridis(np.array([0.0, 1.0])))   # clipped


def test_spectrogram_image_geometry_and_orientation(tmp_path):
    db = np.full((164, 145), -np.inf)
    db[:20, :] = 40.0                       # lowest frequencies hot -> must end up at the BOTTOM (origin='lower')
    img = render.spectrogram_image(db, vmin=10.0, vmax=40.0)
    assert img.size == (496, 370)
    a = np.asarray(img).astype(int)
    top, bottom = a[5, 250], a[-5, 250]
    assert np.abs(top - np.array([68, 1, 84])).max() <= 6           # -inf clips to the bottom colour of the map
    assert np.abs(bottom - np.array([253, 231, 37])).max() <= 6
    p = render.save_spectrogram_jpg(str(tmp_path / "s.jpg"), db, 10.0, 40.0)
    from PIL import Image
    assert Image.open(p).size == (496, 370) and Image.open(p).format == "JPEG"


def test_event_figure_writes_png(tmp_path):
    crop = dict(sxx_db=np.random.default_rng(0).normal(-20, 5, (40, 60)), t=np.linspace(0.1, 8.0, 60),
                f=np.linspace(943, 1063, 40), pxx_db=np.linspace(-30, -10, 80), f_psd=np.linspace(943, 1063, 80),
                t_min=3.0, t_max=4.2)
    p = render.event_figure(str(tmp_path / "e.png"), crop, "Detection from 3.00s to 4.20s\nn_fft: 1024")
    from PIL import Image
    im = Image.open(p)
    assert im.size == (1400, 500) and im.format == "PNG"
    a = np.asarray(im.convert("RGB")).astype(int)
    assert ((a[:, :, 0] > 200) & (a[:, :, 1] < 60) & (a[:, :, 2] < 60)).sum() > 100      # the red marker lines
