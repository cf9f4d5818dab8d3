"""KITTI dataset adapter of the object-detection pipeline — same constructor, item tuple and output files as the
reference's ``tools/datasets.py`` (od/ds:40-190), so ``insertion.py`` finds the data where the reference's script
does.  Differences: no interactive prompt (``create_directories`` takes the run-folder number as an argument and
only asks when ``interactive=True``), and ``save_result`` writes the engine's already compacted float32 records
instead of casting an N x 9 float64 working array (``save_data`` keeps the reference's signature for that case).
"""
import os

import numpy as np


def create_read_me(save_folder, config):
    """``setting.txt`` of a run folder (od/ds:6-17)."""
    ins = config['insertion']
    with open(f'{save_folder}/setting.txt', 'w') as txt:
        txt.write('Inserted classes:\n')
        if ins['random']:
            for c in ins['classes']:
                txt.write('     ' + c + '\n')
            txt.write('Randomly inserted ' + str(ins['number_of_object']) + ' objects\n')
        else:
            for c, n in zip(ins['classes'], ins['number_of_classes']):
                txt.write('     ' + str(n) + 'x   ' + c + '\n')


def create_annotation(old_address, new_address, additional_annotations_lines):
    """Copy of the frame's ``label_2`` file followed by the lines of the inserted objects (od/ds:20-37)."""
    with open(old_address, 'r') as old_txt:
        lines = old_txt.readlines()
    with open(new_address, 'w') as new_txt:
        new_txt.writelines(lines)
        new_txt.writelines(additional_annotations_lines)


def pick_run_folder(root, folder_number=None, interactive=False):
    """Run folders are numbered 00-99 (od/ds:130-175).  Without a number the first free one is NOT chosen silently:
    like the reference, folder 00 is reused unless the caller (or, interactively, the user) says otherwise."""
    if folder_number is None:
        folder_number = 0
        if interactive and os.path.exists(f'{root}/{folder_number:02d}'):
            print('Default save folder is already existing do you want change name? [yes/no]')
            if input() == 'yes':
                print('Write name of the save folder.')
                folder_number = int(input())
    if not 0 <= int(folder_number) <= 99:
        raise ValueError('Input must be number between 0 and 99')
    os.makedirs(f'{root}/{int(folder_number):02d}', exist_ok=True)
    return int(folder_number)


class KITTI():
    def __init__(self, config):
        self.config = config
        self.data_path = config['path']['dataset_path']
        self.label_path = config['path']['label_path']
        self.train_txt_path = config['path']['train_txt_path']
        self.save_output_folder = config['path']['output_path']
        self.velodyne_list = np.array([])
        self.create_velodyne_list()

    def __len__(self):
        return len(self.velodyne_list)

    @staticmethod
    def frame_name(file):
        return file.split('/')[-1].split('.')[0]

    def read_frame(self, idx):
        """(xyzi float32 N x 4, semantic labels uint32 N, frame name): the engine's input layout, no float64 copy."""
        file = self.velodyne_list[idx]
        name = self.frame_name(file)
        xyzi = np.fromfile(file, dtype=np.float32).reshape(-1, 4)
        labels = np.fromfile(f'{self.label_path}/{name}.label', dtype=np.uint32)
        return xyzi, labels & 0xFFFF, name

    def __getitem__(self, idx):
        """The reference's item tuple (od/ds:56-71): N x 5 float64 cloud, annotation path, instance ids, calib, image."""
        xyzi, _, name = self.read_frame(idx)
        labels = np.fromfile(f'{self.label_path}/{name}.label', dtype=np.uint32).reshape(-1, 1)
        pcl = np.hstack((xyzi, labels & 0xFFFF))
        return (pcl, f'{self.data_path}/label_2/{name}.txt', labels >> 16, f'{self.data_path}/calib/{name}.txt',
                f'{self.data_path}/image_2/{name}.png')

    def delete_item(self, idx):
        self.velodyne_list = np.delete(self.velodyne_list, idx)

    def remove_space_for_spherical(self, point_cloud):
        """N x 9 working rows -> N x 4 (x, y, z, intensity) (od/ds:97-109)."""
        pcl = np.ones((len(point_cloud), 4)) * -1
        pcl[:, 0:3] = point_cloud[:, 0:3]
        pcl[:, 3] = point_cloud[:, 6]
        return pcl

    def _write(self, folder, name, velodyne_f32, check_f32, additional_anno_lines):
        out = f'{self.save_output_folder}/{folder}'
        create_annotation(f'{self.data_path}/label_2/{name}.txt', f'{out}/label_2/{name}.txt', additional_anno_lines)
        np.ascontiguousarray(velodyne_f32, dtype=np.float32).tofile(f'{out}/velodyne/{name}.bin')
        np.ascontiguousarray(check_f32, dtype=np.float32).tofile(f'{out}/check/{name}.bin')

    def save_data(self, point_cloud, added_points, folder, name, idx, additional_anno_lines):
        """Reference signature (od/ds:76-95): N x 9 float64 working arrays in."""
        self._write(folder, name, self.remove_space_for_spherical(point_cloud).astype(np.float32),
                    self.remove_space_for_spherical(added_points).astype(np.float32), additional_anno_lines)
        self.delete_item(idx)

    def save_result(self, result, folder, name):
        """Engine result (already float32, compacted on the GPU) -> the same three files."""
        self._write(folder, name, result.velodyne, result.check, result.lines)

    def create_velodyne_list(self):
        with open(self.train_txt_path) as train_txt:
            frames = [int(line) for line in train_txt if len(line.strip()) > 0]
        self.velodyne_list = [f'{self.data_path}/velodyne/{frame:06d}.bin' for frame in frames]

    def create_directories(self, save_folder, folder_number=None, interactive=False):
        root = f"{self.config['path']['output_path']}/{save_folder}"
        os.makedirs(root, exist_ok=True)
        folder_number = pick_run_folder(root, folder_number, interactive)
        save_folder = f'{save_folder}/{folder_number:02d}'
        for sub in ('velodyne', 'check', 'label_2', 'added_objects'):
            os.makedirs(f"{self.config['path']['output_path']}/{save_folder}/{sub}", exist_ok=True)
        create_read_me(f"{self.config['path']['output_path']}/{save_folder}", self.config)
        return save_folder, folder_number
