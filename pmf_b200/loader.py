"""PerspectiveViewLoader with the per-frame projection + scatter done on the B200.

Mirrors ``pc_processor/dataset/perspective_view_loader.py:8-141`` of the reference: same constructor arguments, same
returned tuple (``(feat[8], mask, label)``, or with ``return_uproj`` the six-tuple ``infer.py:78`` unpacks), same
augmentation pipeline (ColorJitter on the PIL image, then RandomHorizontalFlip / RandomRotation / RandomCrop or
CenterCrop and Pad on the stacked 10-channel tensor, drawn from the same torch RNG in the same order).  What changes is
WHERE the frame is built: the reference does ``mapLidar2Camera`` (parser.py:209-227) and five numpy fancy-index scatters
(:92-117) in a DataLoader worker on the host; here the points and mapped labels are copied to the device once and
``pmfb_project_scatter`` produces depth / xyzi / mask / label images in two kernels (bit-exact to the numpy path:
float64 projection, strict FOV test, int truncation, last point wins on a pixel collision).

The item tensors are CUDA tensors.  CUDA cannot be used in forked DataLoader workers, so this loader must run in the
main process: set ``n_threads: 0`` in the task's YAML (the only configuration change; the projection that the worker
processes existed to hide is now ~0.1 ms of device time per frame).
"""
import numpy as np
import torch
from torch.utils.data import Dataset, get_worker_info
from torchvision import transforms

from .postproc import project_scatter


class PerspectiveViewLoader(Dataset):
    def __init__(self, dataset, config, data_len=-1, is_train=True, pcd_aug=False, img_aug=False, use_padding=False,
                 return_uproj=False, device=None):
        self.dataset = dataset
        self.config = config
        self.is_train = is_train
        self.pcd_aug = pcd_aug
        self.img_aug = img_aug
        self.data_len = data_len
        self.use_padding = use_padding
        self.device = device

        if not self.is_train:
            self.pcd_aug = False
            self.img_aug = False
        augment_config = self.config["augmentation"]

        if self.pcd_aug:
            # host-side point-cloud augmentation stays the reference's own code (preprocess/augmentor.py), imported from
            # the package this module was deployed into
            from pc_processor.dataset.preprocess import augmentor
            p = augmentor.AugmentParams()
            p.setFlipProb(p_flipx=augment_config["p_flipx"], p_flipy=augment_config["p_flipy"])
            p.setTranslationParams(
                p_transx=augment_config["p_transx"], trans_xmin=augment_config["trans_xmin"],
                trans_xmax=augment_config["trans_xmax"], p_transy=augment_config["p_transy"],
                trans_ymin=augment_config["trans_ymin"], trans_ymax=augment_config["trans_ymax"],
                p_transz=augment_config["p_transz"], trans_zmin=augment_config["trans_zmin"],
                trans_zmax=augment_config["trans_zmax"])
            p.setRotationParams(
                p_rot_roll=augment_config["p_rot_roll"], rot_rollmin=augment_config["rot_rollmin"],
                rot_rollmax=augment_config["rot_rollmax"], p_rot_pitch=augment_config["p_rot_pitch"],
                rot_pitchmin=augment_config["rot_pitchmin"], rot_pitchmax=augment_config["rot_pitchmax"],
                p_rot_yaw=augment_config["p_rot_yaw"], rot_yawmin=augment_config["rot_yawmin"],
                rot_yawmax=augment_config["rot_yawmax"])
            self.augmentor = augmentor.Augmentor(p)
        else:
            self.augmentor = None

        self.img_jitter = transforms.ColorJitter(*augment_config["img_jitter"]) if self.img_aug else None

        projection_config = self.config["sensor"]
        if self.use_padding:
            h_pad = projection_config["h_pad"]
            w_pad = projection_config["w_pad"]
            self.pad = transforms.Pad((w_pad, h_pad))
        else:
            h_pad = 0
            w_pad = 0
        if self.is_train:
            self.aug_ops = transforms.Compose([
                transforms.RandomHorizontalFlip(0.5),
                transforms.RandomRotation(15),
                transforms.RandomCrop(size=(projection_config["proj_ht"] - 2 * h_pad,
                                            projection_config["proj_wt"] - 2 * w_pad)),
            ])
        else:
            self.aug_ops = transforms.Compose([
                transforms.CenterCrop((projection_config["proj_h"] - 2 * h_pad, projection_config["proj_w"] - 2 * w_pad))
            ])
        self.return_uproj = return_uproj

    def _device(self):
        if get_worker_info() is not None:
            raise RuntimeError("pmf_b200 PerspectiveViewLoader projects on the GPU and cannot run in a DataLoader worker "
                               "process: set n_threads: 0 in the task config")
        if self.device is not None:
            return torch.device(self.device)
        if not torch.cuda.is_available():
            raise RuntimeError("pmf_b200 PerspectiveViewLoader needs a B200 (no CPU fallback; use the reference loader)")
        return torch.device("cuda", torch.cuda.current_device())

    def __getitem__(self, index):
        dev = self._device()
        pointcloud, sem_label, _ = self.dataset.loadDataByIndex(index)
        if self.pcd_aug:
            pointcloud = self.augmentor.doAugmentation(pointcloud)
        image = self.dataset.loadImage(index)
        if self.img_aug:
            image = self.img_jitter(image)
        image = np.array(image)
        seq_id, _ = self.dataset.parsePathInfoByIndex(index)
        h, w = image.shape[0], image.shape[1]

        # device side: projection (parser.py:209-227) + scatter (perspective_view_loader.py:87-117)
        points = torch.from_numpy(np.ascontiguousarray(pointcloud[:, :4], dtype=np.float32)).to(dev, non_blocking=True)
        labels = torch.from_numpy(np.ascontiguousarray(self.dataset.labelMapping(sem_label), dtype=np.int32)).to(
            dev, non_blocking=True)
        proj = project_scatter(points, labels, self.dataset.proj_matrix[seq_id], h, w)
        # uint8 -> [0,1] exactly as perspective_view_loader.py:95 (a true fp32 division; torch's CUDA scalar division
        # multiplies by the reciprocal, which differs in the last bit)
        image_tensor = torch.from_numpy(image.astype(np.float32) / 255.0).to(dev, non_blocking=True).permute(2, 0, 1)
        proj_tensor = torch.cat((proj["feat"], image_tensor, proj["mask"].unsqueeze(0), proj["label"].unsqueeze(0)), dim=0)

        if self.return_uproj:
            keep = proj["keep"]
            # the reference returns int32 row / column indices of the kept points and the depth of ALL points (:133-135)
            return (proj_tensor[:8], proj_tensor[8], proj_tensor[9], proj["rows"][keep], proj["cols"][keep], proj["depth"])
        proj_tensor = self.aug_ops(proj_tensor)
        if self.use_padding:
            proj_tensor = self.pad(proj_tensor)
        return proj_tensor[:8], proj_tensor[8], proj_tensor[9]

    def __len__(self):
        if self.data_len > 0 and self.data_len < len(self.dataset):
            return self.data_len
        return len(self.dataset)
