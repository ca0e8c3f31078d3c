"""PCKh@0.5 of the LIP pose csv files with the counting on the GPU — drop-in for utils/calc_pckh.py."""
import csv

import numpy as np
import torch

from .._lib import call, fptr, i32, f64, stream


def read_data(path, additional_dim):
    """utils/calc_pckh.py:6-33: csv rows `name, x0, y0[, v0], ...`; 'nan' -> -1."""
    labels = []
    with open(path, "r") as f:
        for row in csv.reader(f, delimiter=","):
            labels.append([-1.0 if v == "nan" else float(v) for v in row[1:]])
    data = np.array(labels)
    dim = 3 if additional_dim else 2
    data = np.reshape(data, [data.shape[0], int(data.shape[1] / dim), dim])
    vis_label = np.zeros((data.shape[0], data.shape[1]))
    if additional_dim:
        vis_label[:, :] = data[:, :, 2]
        data = data[:, :, 0:2]
    else:
        vis_label = vis_label + 1
        data[data < 0] = 1
    return data, vis_label


def pckh_counts(pred, gt, thr=0.5):
    """Per-joint (hit, valid) int64 counters of get_head_size / get_norm_dist / compute_pck (:35-97)."""
    p = torch.from_numpy(np.ascontiguousarray(pred, dtype=np.float64)).cuda()
    g = torch.from_numpy(np.ascontiguousarray(gt, dtype=np.float64)).cuda()
    n, j, _ = p.shape
    hit = torch.zeros(j, dtype=torch.int64, device="cuda")
    valid = torch.zeros(j, dtype=torch.int64, device="cuda")
    call("npp_pckh_counts", fptr(p), fptr(g), i32(n), i32(j), f64(thr), fptr(hit), fptr(valid), stream())
    return hit.cpu().numpy(), valid.cpu().numpy()


def pck_from_counts(hit, valid):
    """compute_pck (:58-84): per joint, upper body (8:16) and all (0:6 + 8:16), in percent."""
    P = hit.shape[0]
    pck = np.zeros([1, P + 2])
    with np.errstate(invalid="ignore", divide="ignore"):
        for p in range(P):
            pck[0, p] = 100 * (hit[p] / valid[p]) if valid[p] else np.nan
        ub = list(range(8, 16))
        pck[0, P] = 100 * (hit[ub].sum() / valid[ub].sum()) if valid[ub].sum() else np.nan
        al = list(range(0, 6)) + list(range(8, 16))
        pck[0, P + 1] = 100 * (hit[al].sum() / valid[al].sum()) if valid[al].sum() else np.nan
    return pck


def calc_pck_lip_dataset(gt_path, pred_path, method_name="Ours", eval_num=5000, verbose=True):
    """utils/calc_pckh.py:99-126."""
    pred, _ = read_data(pred_path, False)
    pred = pred[0:eval_num, :, :]
    gt, _ = read_data(gt_path, True)
    gt = gt[0:eval_num, :, :]
    assert gt.shape[0] == pred.shape[0], "sample not matched"
    assert gt.shape[1] == pred.shape[1], "joints not matched"
    assert gt.shape[2] == pred.shape[2], "dim not matched"
    hit, valid = pckh_counts(pred, gt, 0.5)
    pck = pck_from_counts(hit, valid)
    if verbose:
        p = pck[-1]
        tmpl = "{0:10} & {1:6} & {2:6} & {3:6} & {4:6} & {5:6} & {6:6} & {7:6} & {8:6} & {9:6}"
        print(tmpl.format("PCKh@0.5", "Head", "Sho.", "Elb.", "Wri.", "Hip", "Knee", "Ank.", "U.Body", "Avg."))
        print(tmpl.format(method_name, "%1.1f" % ((p[8] + p[9]) / 2.0), "%1.1f" % ((p[12] + p[13]) / 2.0),
                          "%1.1f" % ((p[11] + p[14]) / 2.0), "%1.1f" % ((p[10] + p[15]) / 2.0),
                          "%1.1f" % ((p[2] + p[3]) / 2.0), "%1.1f" % ((p[1] + p[4]) / 2.0),
                          "%1.1f" % ((p[0] + p[5]) / 2.0), "%1.1f" % p[-2], "%1.1f" % p[-1]))
    return pck
