"""Mirror of the reference's src/evaluation.py on libard_b200.so.

train_and_evaluate_residual :19-72, evaluate_zero_shot :74-109, evaluate_baseline_clap :112-130 keep their names, arguments
and the per-fold `.npz` files (similarities / predictions / targets) they write. `eval_metrics` produces the NUMBERS of
visualize_eval_metrics :132-198 (top-1, top-k, macro precision / recall / F1 per fold, mean and std(ddof=1) over folds, the
aggregated confusion matrix) with the counting done on the GPU (`ard_eval_metrics`); the seaborn heat-map is plotting and
out of scope.
"""
import gc
import os

import numpy as np
import torch
import torch.nn as nn

from .head import eval_metrics as _device_metrics
from .head import head_logits
from .residual import quantize_tensor, setup_residual_htsat
from .training import train_one_epoch_zero_shot


def train_and_evaluate_residual(clap, dataset_name, folds, text_embeds, pca_path, save_dir, epochs=10, lr=0.01, inject_layers=[0]):
    """src/evaluation.py:19-72: per fold, inject ResiDual from `pca_path/dataset_name/layer_{l}_evalfold_{i}`, train the lambdas
    with Adam, evaluate zero-shot on the validation loader and save `layers_{..}_evalfold_{i}.npz`.
    Like the reference, every fold patches the encoder it finds on `clap` (a deep copy each time)."""
    device = clap.device
    layers_str = "_".join(map(str, inject_layers))
    save_dir = os.path.join(save_dir, dataset_name, "ResiDual")
    os.makedirs(save_dir, exist_ok=True)
    for i, (train_load, val_load) in enumerate(folds):
        print(f"===== Eval fold {i} =====")
        save_file = os.path.join(save_dir, f"layers_{layers_str}_evalfold_{i}.npz")
        pca_files = {l: os.path.join(pca_path, dataset_name, f"layer_{l}_evalfold_{i}") for l in inject_layers}
        audio_encoder = clap.model.audio_branch
        new_htsat, residuals = setup_residual_htsat(audio_encoder, pca_files, inject_layers)
        clap.model.audio_branch = new_htsat
        optimizer = torch.optim.Adam([res.learnable for res in residuals.values()], lr=lr)
        criterion = nn.CrossEntropyLoss()
        for e in range(epochs):
            print(f"=== Epoch {e} ===")
            train_loss, train_acc = train_one_epoch_zero_shot(clap, train_load, text_embeds, optimizer, criterion, device)
            print(f"Train loss: {train_loss}, Train accuracy: {train_acc}")
        preds, targs, similarities = evaluate_zero_shot(clap, val_load, text_embeds, device)
        np.savez_compressed(save_file, similarities=similarities, predictions=np.array(preds), targets=np.array(targs))
        torch.cuda.empty_cache()
        gc.collect()


def evaluate_zero_shot(model, dataloader, text_embeddings, device):
    """src/evaluation.py:74-109: (predictions, targets, similarities[N, classes]) with int16-quantised inputs (numpy route)."""
    model.eval()
    all_preds, all_targets, all_similarities = [], [], []
    text = text_embeddings.to(device)
    with torch.no_grad():
        for x, true_labels in dataloader:
            audio_data = quantize_tensor(x.squeeze(1)).cpu().numpy()
            audio_embeds = model.get_audio_embedding_from_data(x=audio_data, use_tensor=False)
            audio_embeds = torch.tensor(audio_embeds).to(device).float()
            similarities = head_logits(audio_embeds, text)
            all_preds.extend(similarities.argmax(dim=-1).cpu().tolist())
            all_targets.extend(true_labels.tolist())
            all_similarities.append(similarities.cpu())
    return all_preds, all_targets, torch.cat(all_similarities, dim=0).numpy()


def evaluate_baseline_clap(clap, dataset_name, folds, text_embeds, save_dir):
    """src/evaluation.py:112-130: zero-shot evaluation of the un-patched model on every fold's validation loader."""
    device = clap.device
    save_dir = os.path.join(save_dir, dataset_name, "Baseline")
    os.makedirs(save_dir, exist_ok=True)
    for i, (_, val_load) in enumerate(folds):
        save_file = os.path.join(save_dir, f"evalfold_{i}.npz")
        preds, targs, similarities = evaluate_zero_shot(clap, val_load, text_embeds, device)
        np.savez_compressed(save_file, similarities=similarities, predictions=np.array(preds), targets=np.array(targs))


def fold_metrics(similarities, predictions, targets, n_classes, k_top=5):
    """One fold of src/evaluation.py:149-177 from a device-built confusion matrix: accuracy, top-k accuracy, macro
    precision / recall / F1 with zero_division=0 over the classes present in y_true or y_pred (sklearn's label set)."""
    m = _device_metrics(similarities, targets, k=k_top, n_classes=n_classes)
    pred_cm = np.zeros((n_classes, n_classes), dtype=np.int64)   # the saved predictions define precision/recall (they are the argmax)
    np.add.at(pred_cm, (np.asarray(targets, dtype=np.int64), np.asarray(predictions, dtype=np.int64)), 1)
    if not np.array_equal(pred_cm, m["confusion"]):
        cm = pred_cm          # predictions saved from other scores (e.g. softmax of logits): trust the file
    else:
        cm = m["confusion"]
    tp = np.diag(cm).astype(np.float64)
    support, predicted = cm.sum(axis=1).astype(np.float64), cm.sum(axis=0).astype(np.float64)
    present = (support + predicted) > 0
    prec = np.divide(tp, predicted, out=np.zeros_like(tp), where=predicted > 0)
    rec = np.divide(tp, support, out=np.zeros_like(tp), where=support > 0)
    f1 = np.divide(2 * prec * rec, prec + rec, out=np.zeros_like(tp), where=(prec + rec) > 0)
    n = max(int(cm.sum()), 1)
    return {"acc": tp.sum() / n, "topk": m["topk"], "prec": prec[present].mean(), "rec": rec[present].mean(), "f1": f1[present].mean(),
            "confusion": cm}


def eval_metrics(save_dir, n_classes, n_folds, inject_layers, k_top=5, verbose=True):
    """The numbers of visualize_eval_metrics (src/evaluation.py:132-198) from the per-fold .npz files: per-fold arrays, their
    mean / std(ddof=1), and the confusion matrix summed over folds. `save_dir` is the directory holding the fold files."""
    layers_str = "_".join(map(str, inject_layers)) if inject_layers != [] else ""
    per_fold = {"acc": [], "topk": [], "prec": [], "rec": [], "f1": []}
    agg_cm = np.zeros((n_classes, n_classes), dtype=np.int64)
    for i in range(n_folds):
        name = f"layers_{layers_str}_evalfold_{i}.npz" if layers_str else f"evalfold_{i}.npz"
        data = np.load(os.path.join(save_dir, name))
        fm = fold_metrics(data["similarities"], data["predictions"], data["targets"], n_classes, k_top)
        for k in per_fold:
            per_fold[k].append(fm[k])
        agg_cm += fm["confusion"]
    out = {k: np.asarray(v, dtype=float) for k, v in per_fold.items()}
    summary = {k: (float(v.mean()), float(v.std(ddof=1)) if len(v) > 1 else float("nan")) for k, v in out.items()}
    if verbose:
        print("== Cross-Fold Evaluation Metrics ==")
        print(f"Top-1 Accuracy:   {summary['acc'][0]:.4f} ± {summary['acc'][1]:.4f}")
        print(f"Top-{k_top} Accuracy:  {summary['topk'][0]:.4f} ± {summary['topk'][1]:.4f}")
        print(f"Precision: {summary['prec'][0]:.4f} ± {summary['prec'][1]:.4f}")
        print(f"Recall:    {summary['rec'][0]:.4f} ± {summary['rec'][1]:.4f}")
        print(f"F1:        {summary['f1'][0]:.4f} ± {summary['f1'][1]:.4f}")
    return {"per_fold": out, "summary": summary, "confusion": agg_cm}


def visualize_eval_metrics(save_dir, dataset_name, n_folds, inject_layers, k_top=5, n_classes=None):
    """src/evaluation.py:132-198 without the heat-map: prints the cross-fold metrics and returns them. The reference reads the
    class count from its dataset registry (data_processing.DATASETS, out of scope); pass `n_classes`, or it is taken from the
    width of the saved similarities."""
    if n_classes is None:
        layers_str = "_".join(map(str, inject_layers)) if inject_layers != [] else ""
        name = f"layers_{layers_str}_evalfold_0.npz" if layers_str else "evalfold_0.npz"
        n_classes = int(np.load(os.path.join(save_dir, name))["similarities"].shape[1])
    return eval_metrics(save_dir, n_classes, n_folds, inject_layers, k_top)
