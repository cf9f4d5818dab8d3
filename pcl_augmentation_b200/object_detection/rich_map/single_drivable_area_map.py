"""Rich-map generation of the object-detection pipeline on the GPU — the reference's
``object_detection/rich_map/single_drivable_area_map.py`` (per frame: 1 m road raster, ``closing(disk(4))`` -> road
map, 8-neighbour ring + ``dilation(disk(2))`` -> pedestrian-area map, ``.npz{map, min_x, min_y}``, :123-194).

    python -m pcl_augmentation_b200.object_detection.rich_map.single_drivable_area_map [--config ../config/KITTI.yaml]

walks ``train.txt`` like the reference's script (:76-194) and writes
``<maps_path>/maps/{road_maps,pedestrian_area}/npz/<frame>.npz``; ``drivable_area_maps_batch`` is the batched operator.
"""
import ctypes as C
import os

import numpy as np

from ... import _lib


def drivable_area_maps_batch(xyzi_list, labels_list, road_label):
    """``xyzi_list``: per frame N x 4 float32 (velodyne/*.bin), ``labels_list``: per frame N semantic labels.
    Returns per frame ``(road {'map','min_x','min_y'}, pedestrian {'map','min_x','min_y'})`` with uint8 X x Y maps."""
    import torch
    _lib.require_cuda()
    lib = _lib.load()
    n = len(xyzi_list)
    assert n == len(labels_list) and n > 0
    offs = np.zeros(n + 1, dtype=np.int64)
    for i, p in enumerate(xyzi_list):
        offs[i + 1] = offs[i] + len(p)
    xyzi = torch.from_numpy(np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float32).reshape(-1, 4)
                                                                 for p in xyzi_list]))).cuda()
    labels = torch.from_numpy(np.concatenate([np.asarray(l).reshape(-1).astype(np.uint32).view(np.int32)
                                              for l in labels_list])).cuda()
    d_offs = torch.from_numpy(offs).cuda()
    d_dims = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.r3d_rich_map_od_extents(xyzi.data_ptr(), d_offs.data_ptr(), n, d_dims.data_ptr(), stream), "rich_map_od_extents")
    dims = d_dims.cpu().numpy()
    moff = np.zeros(n + 1, dtype=np.int64)
    moff[1:] = np.cumsum(dims[:, 0].astype(np.int64) * dims[:, 1])
    total = int(moff[-1])
    d_moff = torch.from_numpy(moff).cuda()
    road = torch.empty(max(total, 1), dtype=torch.uint8, device="cuda")
    ped = torch.empty(max(total, 1), dtype=torch.uint8, device="cuda")
    scratch = torch.empty(max(2 * total, 1), dtype=torch.uint8, device="cuda")
    _lib.check(lib.r3d_rich_map_od_build(xyzi.data_ptr(), labels.data_ptr(), d_offs.data_ptr(), n, int(road_label),
                                         d_dims.data_ptr(), d_moff.data_ptr(), total, road.data_ptr(), ped.data_ptr(),
                                         scratch.data_ptr(), stream), "rich_map_od_build")
    h_road, h_ped = road.cpu().numpy(), ped.cpu().numpy()
    out = []
    for i in range(n):
        sx, sy, mx, my = (int(v) for v in dims[i])
        a, b = moff[i], moff[i + 1]
        out.append(({'map': h_road[a:b].reshape(sx, sy).copy(), 'min_x': mx, 'min_y': my},
                    {'map': h_ped[a:b].reshape(sx, sy).copy(), 'min_x': mx, 'min_y': my}))
    return out


def drivable_area_maps(point_cloud, road_label):
    """One frame in the reference's N x 5 layout (x, y, z, intensity, label; ``KITTI.__getitem__``)."""
    pc = np.asarray(point_cloud)
    return drivable_area_maps_batch([pc[:, :4].astype(np.float32)], [pc[:, 4].astype(np.uint32)], road_label)[0]


def generate_maps(config, batch_size=64, log=print):
    """The reference script's loop (:76-194) over the ``train.txt`` frames, ``batch_size`` frames per launch."""
    from ..Real3DAug.tools.datasets import KITTI
    dataset = KITTI(config)
    save_path = config['path']['maps_path']
    for sub in ('road_maps', 'pedestrian_area'):
        os.makedirs(f'{save_path}/maps/{sub}/npz', exist_ok=True)
    road_label = config['labels']['Road']
    done = 0
    for i0 in range(0, len(dataset), batch_size):
        frames = [dataset.read_frame(i) for i in range(i0, min(i0 + batch_size, len(dataset)))]
        maps = drivable_area_maps_batch([f[0] for f in frames], [f[1] for f in frames], road_label)
        for (_, _, name), (road, ped) in zip(frames, maps):
            frame_number = int(name)
            np.savez(f'{save_path}/maps/road_maps/npz/{frame_number:06d}', map=road['map'], min_x=road['min_x'], min_y=road['min_y'])
            np.savez(f'{save_path}/maps/pedestrian_area/npz/{frame_number:06d}', map=ped['map'], min_x=ped['min_x'],
                     min_y=ped['min_y'])
            done += 1
        log(f'{done} / {len(dataset)} frames')
    return done


def main(argv=None):
    import argparse
    import yaml
    ap = argparse.ArgumentParser(description="per-frame road / pedestrian-area maps (KITTI) on the GPU")
    ap.add_argument("--config", default="../config/KITTI.yaml")
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args(argv)
    with open(args.config, "r") as f:
        generate_maps(yaml.safe_load(f), args.batch)


if __name__ == "__main__":
    main()
