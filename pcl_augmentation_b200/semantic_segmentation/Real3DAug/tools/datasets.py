"""SemanticKITTI / Waymo dataset adapters of the semantic-segmentation pipeline — constructor arguments, item tuples
and output files of the reference's ``tools/datasets.py`` (ss/ds:20-215, 218-410).  No interactive prompts: what the
reference asks on stdin (reverse order / scenes to skip, ss/ds:117-141; run folder, ss/ds:143-175) are arguments.
"""
import glob
import os

import numpy as np

from ....object_detection.Real3DAug.tools.datasets import pick_run_folder


def create_read_me(save_folder, config):
    """``setting.txt`` (ss/ds:6-17): class names through ``config['labels']``."""
    ins = config['insertion']
    with open(f'{save_folder}/setting.txt', 'w') as txt:
        txt.write('Inserted classes:\n')
        if ins['random']:
            for c in ins['classes']:
                txt.write('     ' + config['labels'][c] + '\n')
            txt.write('Randomly inserted ' + str(ins['number_of_object']) + ' objects\n')
        else:
            for c, n in zip(ins['classes'], ins['number_of_classes']):
                txt.write('     ' + str(n) + 'x   ' + config['labels'][c] + '\n')


class SemanticKITTI():
    velo_2_cam = np.array([[7.533745e-03, -9.999714e-01, -6.166020e-04, -4.069766e-03],
                           [1.480249e-02, 7.280733e-04, -9.998902e-01, -7.631618e-02],
                           [9.998621e-01, 7.523790e-03, 1.480755e-02, -2.717806e-01],
                           [0, 0, 0, 1]])
    my_calib = np.array([[0, -1, 0, 0], [0, 0, -1, 0], [1, 0, 0, 0], [0, 0, 0, 1]])

    def __init__(self, config, sequence, reverse=False, skip_scenes=0):
        self.sequence = sequence
        self.config = config
        self.data_path = config['path']['dataset_path']
        self.anno_path = config['path']['annotation_path']
        self.poses = np.loadtxt(f'{self.data_path}/sequences/{self.sequence}/poses.txt')
        self.velodyne_list = np.array([])
        self.create_velodyne_list(reverse, skip_scenes)

    def __len__(self):
        return len(self.velodyne_list)

    @staticmethod
    def frame_name(file):
        return file.split('/')[-1].split('.')[0]

    def read_frame(self, idx):
        """(xyzi float32, semantic labels uint32, lidar->world 4 x 4, bbox annotation path, frame name)."""
        file = self.velodyne_list[idx]
        name = self.frame_name(file)
        xyzi = np.fromfile(file, dtype=np.float32).reshape(-1, 4)
        labels = np.fromfile(f'{self.data_path}/sequences/{self.sequence}/labels/{name}.label', dtype=np.uint32)
        return (xyzi, labels & 0xFFFF, self.create_transform_matrix(self.poses, int(name)),
                f'{self.anno_path}/sequences/{self.sequence}/bbox/{name}.txt', name)

    def __getitem__(self, idx):
        """The reference's item tuple (ss/ds:45-62)."""
        xyzi, _, transform_matrix, anno, name = self.read_frame(idx)
        labels = np.fromfile(f'{self.data_path}/sequences/{self.sequence}/labels/{name}.label',
                             dtype=np.uint32).reshape(-1, 1)
        return np.hstack((xyzi, labels & 0xFFFF)), transform_matrix, anno, labels >> 16, self.sequence

    def delete_item(self, idx):
        self.velodyne_list = np.delete(self.velodyne_list, idx)

    def create_transform_matrix(self, poses, frame_number):
        """lidar -> world of a frame from the KITTI odometry pose row (ss/ds:65-70)."""
        pose = np.vstack((poses[frame_number].reshape(3, 4), np.array([0, 0, 0, 1])))
        return np.dot(np.linalg.inv(self.my_calib), np.dot(pose, self.velo_2_cam))

    def remove_space_for_spherical(self, point_cloud):
        """N x 9 working rows -> (N x 4 x y z intensity, N x 1 labels) (ss/ds:93-106)."""
        pcl = np.ones((len(point_cloud), 4)) * -1
        pcl[:, 0:3] = point_cloud[:, 0:3]
        pcl[:, 3] = point_cloud[:, 6]
        return pcl, point_cloud[:, 7:8].copy()

    def _write(self, folder, name, velodyne_f32, labels_u32, check_f32):
        out = f"{self.config['path']['output_path']}/{folder}"
        np.ascontiguousarray(velodyne_f32, dtype=np.float32).tofile(f'{out}/velodyne/{name}.bin')
        np.ascontiguousarray(labels_u32, dtype=np.uint32).tofile(f'{out}/labels/{name}.label')
        np.ascontiguousarray(check_f32, dtype=np.float32).tofile(f'{out}/check/{name}.bin')

    def save_data(self, point_cloud, added_points, folder, name, idx):
        """Reference signature (ss/ds:72-91): N x 9 float64 working arrays in."""
        point_cloud, pcl_labels = self.remove_space_for_spherical(point_cloud)
        added_points, added_labels = self.remove_space_for_spherical(added_points)
        self._write(folder, name, point_cloud.astype(np.float32), pcl_labels.astype(np.uint32),
                    np.hstack((added_points, added_labels)).astype(np.float32))
        self.delete_item(idx)

    def save_result(self, result, folder, name):
        self._write(folder, name, result.velodyne, result.labels, result.check)

    def create_velodyne_list(self, reverse=False, skip_scenes=0):
        files = sorted(glob.glob(f'{self.data_path}/sequences/{self.sequence}/velodyne/*.bin'))[skip_scenes:]
        self.velodyne_list = np.array(files[::-1] if reverse else files)

    def sequence_folders(self):
        return [f'{s:02d}' for s in self.config['split']['train']]

    def create_directories(self, save_folder, folder_number=None, interactive=False):
        out = self.config['path']['output_path']
        os.makedirs(f'{out}/{save_folder}', exist_ok=True)
        folder_number = pick_run_folder(f'{out}/{save_folder}', folder_number, interactive)
        save_folder = f'{save_folder}/{folder_number:02d}/sequences'
        os.makedirs(f'{out}/{save_folder}', exist_ok=True)
        create_read_me(f'{out}/{save_folder}', self.config)
        for s in self.sequence_folders():
            for sub in ('velodyne', 'check', 'labels', 'added_objects'):
                os.makedirs(f'{out}/{save_folder}/{s}/{sub}', exist_ok=True)
        return save_folder, folder_number


class Waymo(SemanticKITTI):
    """Waymo frames in the reference's converted layout (ss/ds:218-410): ``<data>/<sequence>/lidar/<f>.npy`` (N x 6,
    the first four columns are x y z intensity), ``labels_v3_2/<f>.npy`` (N x 2: instance, semantic),
    ``poses/<f>.npy`` (4 x 4).  Points are shifted by the LiDAR mounting position on the way in and back on the way
    out (ss/ds:225, 262-267, 286-287); outputs go to ``lidar / labels_v3_2 / check`` as ``.npy``.
    Deviation: the engine keeps original points as float32, so the shifted coordinates are rounded to float32 here
    (<= 4e-6 m at 60 m) where the reference carries them in float64."""

    LiDAR_location = np.array([1.22, 0, 2])

    def __init__(self, config, reverse=False, skip_scenes=0):
        self.config = config
        self.data_path = config['path']['dataset_path']
        self.anno_path = config['path']['annotation_path']
        self.velodyne_list = np.array([])
        self.sequence = None
        self.save_subfolder = None
        self.sequence_names = []
        self.create_velodyne_list(reverse, skip_scenes)

    @staticmethod
    def _sibling(pcl_file, folder):
        parts = pcl_file.split('/')
        parts[-2] = folder
        return '/'.join(parts)

    def read_frame(self, idx):
        pcl_file = self.velodyne_list[idx]
        sequence, name = pcl_file.split('/')[-3], self.frame_name(pcl_file)
        pcl = np.load(pcl_file).reshape(-1, 6)[:, :4].astype(np.float64)
        pcl[:, 0:3] -= self.LiDAR_location
        labels = np.load(self._sibling(pcl_file, 'labels_v3_2')).reshape(-1, 2)[:, 1].astype(np.uint32)
        correction = np.eye(4)
        correction[0:3, 3] = self.LiDAR_location
        pose = np.load(self._sibling(pcl_file, 'poses')).reshape(4, 4) @ correction
        return pcl.astype(np.float32), labels, pose, f'{self.anno_path}/{sequence}/bbox/{name}.txt', name

    def __getitem__(self, idx):
        pcl_file = self.velodyne_list[idx]
        xyzi, labels, pose, anno, _ = self.read_frame(idx)
        instances = np.load(self._sibling(pcl_file, 'labels_v3_2')).reshape(-1, 2)[:, 0:1]
        return np.hstack((xyzi, labels.reshape(-1, 1))), pose, anno, instances, pcl_file.split('/')[-3]

    def delete_item(self, idx, subdirectoties=True):
        self.velodyne_list = np.delete(self.velodyne_list, idx)
        if len(self.velodyne_list) == 0:
            self.sequence = None
            return
        sequence = self.velodyne_list[0].split('/')[-3]
        if self.sequence != sequence and subdirectoties:
            self.create_subdirectories(sequence)
        self.sequence = sequence

    def _write(self, folder, name, velodyne_f32, labels_u32, check_f32):
        out = f"{self.config['path']['output_path']}/{folder}"
        cloud = np.array(velodyne_f32, dtype=np.float64)
        cloud[:, 0:3] += self.LiDAR_location
        added = np.array(check_f32, dtype=np.float64)
        if len(added):
            added[:, 0:3] += self.LiDAR_location
        np.save(f'{out}/lidar/{name}.npy', cloud.astype(np.float32))
        np.save(f'{out}/labels_v3_2/{name}.npy', np.asarray(labels_u32, dtype=np.uint32).reshape(-1, 1))
        np.save(f'{out}/check/{name}.npy', added.astype(np.float32))

    def create_velodyne_list(self, reverse=False, skip_scenes=0):
        files = []
        for sequence in sorted(glob.glob(f'{self.data_path}/*/')):
            files += sorted(glob.glob(f'{sequence}lidar/*.npy'))
            self.sequence_names.append(sequence.split('/')[-2])
        files = files[skip_scenes:]
        self.velodyne_list = np.array(files[::-1] if reverse else files)
        self.sequence = self.velodyne_list[0].split('/')[-3] if len(files) else None

    def create_directories(self, save_folder, folder_number=None, interactive=False):
        out = self.config['path']['output_path']
        os.makedirs(f'{out}/{save_folder}', exist_ok=True)
        folder_number = pick_run_folder(f'{out}/{save_folder}', folder_number, interactive)
        self.save_subfolder = f'{save_folder}/{folder_number:02d}'
        create_read_me(f'{out}/{save_folder}', self.config)
        self.create_subdirectories(self.sequence)
        return self.save_subfolder, folder_number

    def create_subdirectories(self, sequence):
        out = f"{self.config['path']['output_path']}/{self.save_subfolder}/{sequence}"
        for sub in ('lidar', 'check', 'labels_v3_2', 'added_objects'):
            os.makedirs(f'{out}/{sub}', exist_ok=True)
