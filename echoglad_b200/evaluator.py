"""Drop-in for the reference's `LandmarkExpectedCoordiantesEvaluator` (sic; `EVALUATORS['landmarkcoorderror']`,
src/builders/evaluator_builder.py:6-12, src/core/evaluators.py:239-483): same constructor, `update / compute /
reset / get_last / get_predictions / get_sum_of_width_*` contract, but the per-node work — softmax over the
224 x 224 main-level logits of every frame and channel, its expected (h, w), the arg-max of the label heat map —
runs in ONE device kernel (`eg_expected_coords`) on the logits the model just produced, so a step transfers
3 x [B,4,2] numbers instead of the `[B*N,4]` logits the reference moves to the host (`src/engine.py:466-492`).
The remaining arithmetic is on [B,4] tensors and follows the reference line by line."""
from __future__ import annotations

import numpy as np
import torch

from ._lib import EchogladError, check, lib

_NAMES = ('lvid_top', 'lvid_bot', 'lvpw', 'ivs')


def expected_coords(y_pred: torch.Tensor, y_true: torch.Tensor, valid, batch_size: int, frame_size: int):
    """-> preds float[B,4,2], gt int32[B,4,2], valid_subset float[B,4] (device tensors)."""
    if not y_pred.is_cuda:
        # the unmodified engine hands the evaluator `.cpu()` copies (src/engine.py:471-490): the kernel still runs
        # on the GPU -- the tensors go back to the current CUDA device (there is no CPU implementation)
        if not torch.cuda.is_available():
            raise EchogladError("expected_coords needs a CUDA device: echoglad_b200 has no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device())
        y_pred, y_true = y_pred.to(dev), y_true.to(dev)
        valid = None if valid is None else valid.to(dev)
    elif y_true.device != y_pred.device or (valid is not None and valid.device != y_pred.device):
        y_true = y_true.to(y_pred.device)
        valid = None if valid is None else valid.to(y_pred.device)
    lg = y_pred.detach().to(torch.float32).contiguous().view(-1, 4)
    yt = y_true.detach().to(torch.float32).contiguous().view(-1, 4)
    vd = None if valid is None else valid.detach().to(torch.float32).contiguous().view(-1, 4)
    if lg.shape[0] % batch_size or yt.shape != lg.shape or (vd is not None and vd.shape != lg.shape):
        raise EchogladError(f"expected_coords: logits {tuple(lg.shape)} / labels {tuple(yt.shape)} do not match "
                            f"batch_size={batch_size}")
    n0 = lg.shape[0] // batch_size
    dev = lg.device
    preds = torch.empty(batch_size, 4, 2, device=dev)
    gt = torch.empty(batch_size, 4, 2, device=dev, dtype=torch.int32)
    vs = torch.empty(batch_size, 4, device=dev)
    with torch.cuda.device(dev):
        check(lib.eg_expected_coords(batch_size, 4, n0, frame_size, lg.data_ptr(), yt.data_ptr(),
                                     None if vd is None else vd.data_ptr(), preds.data_ptr(), gt.data_ptr(),
                                     vs.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "eg_expected_coords")
    return preds, gt, vs


class LandmarkExpectedCoordiantesEvaluator(object):
    """Locates the landmarks as the expected value of the soft-maxed heat map and measures how far they are from the
    ground truth (src/core/evaluators.py:239-483)."""

    def __init__(self, logger=None, batch_size=2, frame_size=224, use_coord_graph=False):
        if use_coord_graph:
            raise NotImplementedError("the reference evaluator itself raises NameError with use_coord_graph=True "
                                      "(valid_subset undefined, src/core/evaluators.py:302-309,356); heat-map mode only")
        self.batch_size = batch_size
        self.frame_size = frame_size
        self.use_coord_graph = use_coord_graph
        self.detailed_performance = {}
        self.reset()

    def reset(self):
        self.coordinate_errors = {k: [] for k in ('ivs', 'lvid_top', 'lvid_bot', 'lvpw')}
        self.valid_errors = {k: [] for k in ('ivs', 'lvid_top', 'lvid_bot', 'lvpw')}
        self.width_MAE = {k: [] for k in ('lvid', 'ivs', 'lvpw')}
        self.width_MPE = {k: [] for k in ('lvid', 'ivs', 'lvpw')}
        self.detailed_performance.clear()

    @staticmethod
    def get_pixel_length(x0, y0, x1, y1, pix2mm_x, pix2mm_y):
        return torch.sqrt(((x0 - x1) * pix2mm_x) ** 2 + ((y0 - y1) * pix2mm_y) ** 2)

    def update(self, y_pred, y_true, pix2mm_x, pix2mm_y, valid):
        self.detailed_performance.clear()
        preds, gt, valid_subset = expected_coords(y_pred, y_true, valid, self.batch_size, self.frame_size)
        # everything below works on [B,4] numbers; one small D2H transfer, then the reference's arithmetic
        preds, gt, valid_subset = preds.cpu(), gt.cpu().to(torch.int64), valid_subset.cpu()
        pix2mm_x, pix2mm_y = pix2mm_x.detach().cpu().float(), pix2mm_y.detach().cpu().float()
        num_valid = torch.sum(valid_subset, dim=0, keepdim=True)
        for k, name in enumerate(_NAMES):
            self.valid_errors[name].append((num_valid[0, k] > 0).item())
        num_valid[num_valid == 0] = 1
        gt_h, gt_w = gt[:, :, 0], gt[:, :, 1]
        preds_h, preds_w = preds[:, :, 0], preds[:, :, 1]
        err = self.get_pixel_length(gt_w, gt_h, preds_w, preds_h, pix2mm_x.unsqueeze(1), pix2mm_y.unsqueeze(1)).numpy()
        err *= valid_subset.numpy()
        err = np.squeeze(np.sum(err, axis=0) / num_valid.numpy())
        for k, name in enumerate(_NAMES):
            self.coordinate_errors[name].append(err[k])
        widths = self.calculate_widths(preds, gt, pix2mm_x, pix2mm_y)
        scale = {'lvid': valid_subset[:, 0] * valid_subset[:, 1] / torch.min(num_valid[0, 0], num_valid[0, 1]),
                 'ivs': valid_subset[:, 3] / num_valid[0, 3], 'lvpw': valid_subset[:, 2] / num_valid[0, 2]}
        ivs, lvid, lvpw = self.calculate_width_MAE(widths)
        for k, e in (('ivs', ivs), ('lvid', lvid), ('lvpw', lvpw)):
            self.width_MAE[k].append((e * scale[k]).sum().item())
        ivs, lvid, lvpw = self.calculate_width_MPE(widths)
        for k, e in (('ivs', ivs), ('lvid', lvid), ('lvpw', lvpw)):
            self.width_MPE[k].append((e * scale[k]).sum().item())
        coordinates = {'pred_ivs': preds[:, 3], 'pred_lvid_top': preds[:, 0], 'pred_lvid_bot': preds[:, 1],
                       'pred_lvpw': preds[:, 2], 'gt_ivs': gt[:, 3], 'gt_lvid_top': gt[:, 0],
                       'gt_lvid_bot': gt[:, 1], 'gt_lvpw': gt[:, 2]}
        self.detailed_performance = {'widths': widths, 'coordinates': coordinates}

    # width name -> the two landmarks it spans (indices into lvid_top, lvid_bot, lvpw, ivs)
    _SPANS = {'ivs': (3, 0), 'lvid': (0, 1), 'lvpw': (1, 2)}

    def calculate_widths(self, preds, gt, pix2mm_x, pix2mm_y):
        """[B,4,2] coordinates -> {'pred_<w>_mm', 'gt_<w>_mm'}: physical lengths of the three measured widths."""
        out = {}
        for tag, pts in (('pred', preds), ('gt', gt.to(preds.dtype))):
            for name, (a, b) in self._SPANS.items():
                out[f'{tag}_{name}_mm'] = self.get_pixel_length(pts[:, a, 1], pts[:, a, 0], pts[:, b, 1], pts[:, b, 0],
                                                               pix2mm_x, pix2mm_y)
        return out

    @staticmethod
    def _width_abs_err(widths):
        return {k: torch.abs(widths[f'pred_{k}_mm'] - widths[f'gt_{k}_mm']) for k in ('ivs', 'lvid', 'lvpw')}

    def calculate_width_MAE(self, widths):
        e = self._width_abs_err(widths)
        return e['ivs'], e['lvid'], e['lvpw']

    def calculate_width_MPE(self, widths):
        e = self._width_abs_err(widths)
        return tuple(100 * e[k] / widths[f'gt_{k}_mm'] for k in ('ivs', 'lvid', 'lvpw'))

    def compute(self):
        """Means over the recorded batches, each normalised by the number of batches in which the landmark (for
        lvid: both of its landmarks) was valid."""
        ok = {k: np.asarray(v) for k, v in self.valid_errors.items()}
        ok['lvid'] = np.logical_and(ok['lvid_top'], ok['lvid_bot'])
        out = {k: np.asarray(self.coordinate_errors[k]).sum() / np.count_nonzero(ok[k]) for k in _NAMES}
        for suffix, rec in (('_w', self.width_MAE), ('_mpe', self.width_MPE)):
            for k in ('ivs', 'lvid', 'lvpw'):
                out[k + suffix] = np.asarray(rec[k]).sum() / np.count_nonzero(ok[k])
        return out

    def get_sum_of_width_MAE(self):
        temp = self.compute()
        return sum(value for k, value in temp.items() if k in ['ivs_w', 'lvid_w', 'lvpw_w'])

    def get_sum_of_width_MPE(self):
        temp = self.compute()
        return sum(value for k, value in temp.items() if k in ['ivs_mpe', 'lvid_mpe', 'lvpw_mpe'])

    def get_last(self):
        temp = {k: self.coordinate_errors[k][-1] for k in _NAMES}
        for k in ('ivs', 'lvid', 'lvpw'):
            temp[k + '_w'] = self.width_MAE[k][-1]
        for k in ('ivs', 'lvid', 'lvpw'):
            temp[k + '_mpe'] = self.width_MPE[k][-1]
        return temp

    def get_predictions(self):
        return self.detailed_performance

    # ---- visualisation helpers used by Engine.log_heatmap_wandb (src/engine.py:560-578); host-side, not on the hot
    # path: plain torch on whatever device the tensors live on, matplotlib imported only when a figure is asked for
    _CHANNEL_RGB = ((0.0, 1.0, 1.0), (1.0, 0.7, 0.9), (0.0, 1.0, 0.0), (1.0, 0.0, 0.0))

    def get_softmaxed_heatmap(self, y_pred):
        """[N, nodes, C] logits -> [N, S, S, C]: softmax over the main-level nodes of every frame and channel
        (src/core/evaluators.py:451-461)."""
        s = self.frame_size
        main = y_pred[:, -s * s:, :]
        return torch.softmax(main, dim=1).view(-1, s, s, y_pred.shape[-1])

    def create_overlay_image(self, x, hms):
        """Gray frame [S,S] + C heat maps [S,S,C] -> PIL image; overlapping channels keep the brightest colour
        (src/core/evaluators.py:592-613)."""
        import torchvision
        hms = torch.as_tensor(hms, dtype=torch.float32)
        img = (0.8 * torch.as_tensor(x, dtype=torch.float32)).clamp_min(0).expand(3, -1, -1).clone()
        for c, rgb in enumerate(self._CHANNEL_RGB[:hms.shape[-1]]):
            img = torch.maximum(img, torch.tensor(rgb).view(3, 1, 1) * hms[:, :, c])
        return torchvision.transforms.ToPILImage()(img)

    def get_heatmaps(self, x, landmark_preds, landmark_y, coord_preds, pix2mm_x, pix2mm_y):
        """matplotlib figure of the first frame of the batch: predicted vs ground-truth heat maps with the landmark
        points, the three measured widths and their errors (src/core/evaluators.py:463-590).  Needs `update()` to have
        run on the same batch."""
        import matplotlib.pyplot as plt
        s, b = self.frame_size, self.batch_size
        frame = x[0, 0].detach().cpu()
        gt_map = landmark_y.detach().cpu().view(b, -1, 4)[0, -s * s:, :].view(s, s, 4)
        hms = self.get_softmaxed_heatmap(landmark_preds.detach().cpu().view(b, -1, 4)[0:1])[0]
        hms = hms / hms.amax(dim=(0, 1), keepdim=True)
        panels = {'pred': self.create_overlay_image(frame, hms), 'gt': self.create_overlay_image(frame, gt_map)}
        widths = {k: v[0] for k, v in self.detailed_performance['widths'].items()}
        mae = dict(zip(('ivs', 'lvid', 'lvpw'), self.calculate_width_MAE(widths)))
        mpe = dict(zip(('ivs', 'lvid', 'lvpw'), self.calculate_width_MPE(widths)))
        pts = {k: v[0].tolist() for k, v in self.detailed_performance['coordinates'].items()}
        order = ['pred', 'gt']
        if coord_preds is not None:
            cp = coord_preds.detach().cpu().view(b, -1, 2)[0]
            for name, k in zip(_NAMES, range(4)):
                pts[f'coord_{name}'] = cp[k].tolist()
            for name, (i, j) in self._SPANS.items():
                widths[f'coord_{name}_mm'] = self.get_pixel_length(cp[i, 0], cp[i, 1], cp[j, 0], cp[j, 1],
                                                                   pix2mm_x[0], pix2mm_y[0])
            panels['coord'] = self.create_overlay_image(frame, torch.zeros_like(hms))
            order = ['pred', 'coord', 'gt']
        fig, axs = plt.subplots(1, len(order), figsize=(4 * len(order), 5))
        for ax, tag in zip(axs, order):
            ax.imshow(panels[tag])
            for name in _NAMES:
                ax.plot(pts[f'{tag}_{name}'][1], pts[f'{tag}_{name}'][0], marker='x', color='white', markersize=4)
            for a, c, colour in (('ivs', 'lvid_top', 'dodgerblue'), ('lvid_bot', 'lvpw', 'dodgerblue'),
                                 ('lvid_top', 'lvid_bot', 'red')):
                ax.plot([pts[f'{tag}_{a}'][1], pts[f'{tag}_{c}'][1]], [pts[f'{tag}_{a}'][0], pts[f'{tag}_{c}'][0]],
                        color=colour, linewidth=1.5)
            ax.set_title(f"{tag}: [{float(widths[f'{tag}_ivs_mm']):.1f}, {float(widths[f'{tag}_lvid_mm']):.1f}, "
                         f"{float(widths[f'{tag}_lvpw_mm']):.1f}]")
        plt.setp(axs[1].get_yticklabels(), visible=False)
        fig.tight_layout()
        fig.suptitle(f"MAE [mm] | IVS: {float(mae['ivs']):.1f} | LVID: {float(mae['lvid']):.1f} | LVPW: {float(mae['lvpw']):.1f}\n"
                     f"MPE [%] | IVS: {float(mpe['ivs']):.1f} | LVID: {float(mpe['lvid']):.1f} | LVPW: {float(mpe['lvpw']):.1f}")
        return fig
