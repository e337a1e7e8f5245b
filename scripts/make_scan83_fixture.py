"""Build-container script: the 64 DTU scan83 cameras the reference ships (UV-Mapping/data/DTU/scan83/trainData/in_cam*.npy,
read by DtuDataset.initialize, data/dtu.py:64-72) as one small fixture, plus ray directions of view 33 (the reference's
centre camera, dtu.py:113-114) computed by the reference's own get_rays_dir (dtu.py:27-37) at a few pixels, to pin
ngf_b200.synth.scan83_camera."""
import importlib.util, os, sys, types
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = "/root/reference/UV-Mapping/data/DTU/scan83/trainData/"
cams = {k: np.load(D + f"in_cam{k}.npy") for k in ("Orgs", "Extrinsics", "Focal", "Princpt")}
for missing in ("h5py", "trimesh"):
    sys.modules.setdefault(missing, types.ModuleType(missing))
sys.path.insert(0, "/root/reference/UV-Mapping")
spec = importlib.util.spec_from_file_location("_ref_dtu", "/root/reference/UV-Mapping/data/dtu.py")
dtu = importlib.util.module_from_spec(spec)
spec.loader.exec_module(dtu)
v = 33
px, py = np.meshgrid(np.arange(800).astype(np.float32), np.arange(600).astype(np.float32))        # dtu.py:160-163 (no_crop)
pix = np.stack((px, py), axis=-1).astype(np.float32)
dirs = dtu.get_rays_dir(pix, 600, 800, cams["Focal"][v], cams["Extrinsics"][v][0:3, 0:3], cams["Princpt"][v]).reshape(-1, 3)
idx = np.arange(0, 480000, 7919)[:48]
np.savez_compressed(os.path.join(ROOT, "neural-gauge-fields_b200", "data", "scan83_cameras.npz"),
                    campos=cams["Orgs"].astype(np.float64), extrinsics=cams["Extrinsics"].astype(np.float32),
                    focal=cams["Focal"].astype(np.float32), princpt=cams["Princpt"].astype(np.float32),
                    check_view=np.int64(v), check_index=idx, check_raydir=dirs[idx].astype(np.float32))
print("wrote", dirs.shape, dirs.dtype)
