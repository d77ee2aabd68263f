"""Mirror of the reference's src/training.py on libard_b200.so: train_one_epoch_zero_shot :12-41, evaluate :44-69,
train_with_config :72-142 (same model / optimizer wiring; W&B is optional, see `train_with_config`).

The encoder forward/backward, the similarity logits, the cross entropy and their gradients all run in the library
(`ard_encoder_forward/backward`, `ard_head_forward/backward`, `ard_ce_forward`); torch holds the tensors and runs Adam.
"""
import gc
import os

import torch
import torch.nn as nn

from .head import apply_criterion, head_logits
from .residual import quantize_tensor, setup_residual_htsat  # noqa: F401  (re-exported like the reference's `from src import ...`)


def train_one_epoch_zero_shot(model, dataloader, text_embeddings, optimizer, criterion, device):
    """src/training.py:12-41. The encoder runs in eval mode (hook.py:173) while lambda receives gradients; similarities are
    NOT multiplied by a logit scale."""
    model.train()
    total_loss, correct, total = 0.0, 0, 0
    text = text_embeddings.to(device)
    for x, true_labels in dataloader:
        optimizer.zero_grad()
        audio_data = x.squeeze(1).to(device)
        audio_embeds = model.get_audio_embedding_from_data(x=audio_data, use_tensor=True)
        audio_embeds = audio_embeds.to(device).float()
        similarities = head_logits(audio_embeds, text)                       # audio_embeds @ text_embeddings.T
        loss = apply_criterion(criterion, similarities, true_labels.to(device))
        loss.backward()
        optimizer.step()
        preds = similarities.argmax(dim=-1).cpu()
        correct += (preds == true_labels).sum().item()
        total += x.size(0)
        total_loss += loss.item() * x.size(0)
    return total_loss / total, correct / total


def evaluate(model, dataloader, text_embeddings, criterion, device):
    """src/training.py:44-69: int16-quantised inputs through the numpy route (hook.py:177-179)."""
    model.eval()
    total_loss, correct, total = 0.0, 0, 0
    text = text_embeddings.to(device)
    with torch.no_grad():
        for x, true_labels in dataloader:
            audio_data = quantize_tensor(x.squeeze(1)).cpu().numpy()
            audio_embeds = model.get_audio_embedding_from_data(x=audio_data, use_tensor=False)
            audio_embeds = torch.tensor(audio_embeds).to(device).float()
            similarities = head_logits(audio_embeds, text)
            loss = apply_criterion(criterion, similarities, true_labels.to(device))
            preds = similarities.argmax(dim=-1).cpu()
            correct += (preds == true_labels).sum().item()
            total += x.size(0)
            total_loss += loss.item() * x.size(0)
    return total_loss / total, correct / total


class _Config(dict):
    """W&B-style sweep config: attribute access over a dict (config.learning_rate, ...)."""
    __getattr__ = dict.__getitem__


def train_with_config(config, clap, dataset_name, folds, text_embeds, pca_path, project_name="residual-clap", logger=None):
    """src/training.py:72-142: one sweep run = ResiDual on `config.inject_layers`, Adam(lr) over the `learnable` vectors for
    `config.epochs` epochs on fold `config.eval_fold`. The reference reports to Weights & Biases; here `logger`, if given, is
    called with the same dict per epoch (`wandb.log` fits), and the history + best accuracy + final lambdas are returned - the
    W&B service itself is host-side reporting and out of scope. `config` may be a dict or any object with the four attributes."""
    if isinstance(config, dict):
        config = _Config(config)
    lr, epochs, layers, eval_fold = config.learning_rate, config.epochs, list(config.inject_layers), config.eval_fold
    layers_str = "_".join(map(str, layers))
    device = clap.device
    run_name = f"lr={lr}_ep={epochs}_L={layers_str}_evalfold={eval_fold}"
    train_loader, val_loader = folds[eval_fold]
    pca_files = {l: os.path.join(pca_path, dataset_name, f"layer_{l}_evalfold_{eval_fold}") for l in layers}
    audio_encoder = clap.model.audio_branch
    new_htsat, residuals = setup_residual_htsat(audio_encoder, pca_files, layers)
    clap.model.audio_branch = new_htsat
    optimizer = torch.optim.Adam([res.learnable for res in residuals.values()], lr=lr)
    criterion = nn.CrossEntropyLoss()
    best_acc, history = 0.0, []
    for epoch in range(epochs):
        train_loss, train_acc = train_one_epoch_zero_shot(clap, train_loader, text_embeds, optimizer, criterion, device)
        val_loss, val_acc = evaluate(clap, val_loader, text_embeds, criterion, device)
        best_acc = max(best_acc, val_acc)
        rec = {"fold": eval_fold, "epoch": epoch + 1, "train/loss": train_loss, "train/accuracy": train_acc, "val/loss": val_loss,
               "val/accuracy": val_acc}
        history.append(rec)
        if logger is not None:
            logger(rec)
    print(f"Fold {eval_fold} - Best Val Acc: {best_acc:.4f}")
    torch.cuda.empty_cache()
    gc.collect()
    return {"run_name": run_name, "project": project_name, "history": history, "best_val_accuracy": best_acc, "residuals": residuals,
            "final_learnable": {l: r.learnable.detach().cpu().numpy() for l, r in residuals.items()}}


def train_residual(clap, train_loader, val_loader, text_embeds, pca_files, layers, lr=0.01, epochs=10, device=None):
    """Body of train_with_config for explicit loaders / PCA files (used by the tests and the bench). Returns (residuals, history)."""
    device = device or clap.device
    new_htsat, residuals = setup_residual_htsat(clap.model.audio_branch, pca_files, layers)
    clap.model.audio_branch = new_htsat
    optimizer = torch.optim.Adam([res.learnable for res in residuals.values()], lr=lr)
    criterion = nn.CrossEntropyLoss()
    history = []
    for epoch in range(epochs):
        tl, ta = train_one_epoch_zero_shot(clap, train_loader, text_embeds, optimizer, criterion, device)
        vl, va = evaluate(clap, val_loader, text_embeds, criterion, device)
        history.append({"epoch": epoch + 1, "train/loss": tl, "train/accuracy": ta, "val/loss": vl, "val/accuracy": va})
    return residuals, history
