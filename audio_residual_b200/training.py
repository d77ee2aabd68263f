"""Mirror of the reference's src/training.py (train_one_epoch_zero_shot :12-41, evaluate :44-69) on libard_b200.so.

W&B logging of train_with_config (:72-142) is host-side reporting and out of scope; the sweep body is kept as
`train_residual` with the same model/optimizer wiring (Adam over the ResiDual `learnable` leaves, src/training.py:106).
"""
import os

import torch
import torch.nn as nn

from .residual import quantize_tensor, setup_residual_htsat


def train_one_epoch_zero_shot(model, dataloader, text_embeddings, optimizer, criterion, device):
    """src/training.py:12-41. The encoder runs in eval mode (hook.py:173) while lambda receives gradients; similarities are
    NOT multiplied by a logit scale."""
    model.train()
    total_loss, correct, total = 0.0, 0, 0
    for x, true_labels in dataloader:
        optimizer.zero_grad()
        audio_data = x.squeeze(1).to(device)
        audio_embeds = model.get_audio_embedding_from_data(x=audio_data, use_tensor=True)
        audio_embeds = audio_embeds.to(device).float()
        similarities = torch.matmul(audio_embeds, text_embeddings.T.to(device))
        loss = criterion(similarities, true_labels.to(device))
        loss.backward()
        optimizer.step()
        preds = similarities.argmax(dim=-1).cpu()
        correct += (preds == true_labels).sum().item()
        total += x.size(0)
        total_loss += loss.item() * x.size(0)
    return total_loss / total, correct / total


def evaluate(model, dataloader, text_embeddings, criterion, device):
    """src/training.py:44-69: int16-quantised inputs (the quantisation runs on the device inside the encoder call)."""
    model.eval()
    total_loss, correct, total = 0.0, 0, 0
    with torch.no_grad():
        for x, true_labels in dataloader:
            audio_embeds = model.get_audio_embedding_from_data(x=x.squeeze(1), use_tensor=False)
            audio_embeds = torch.as_tensor(audio_embeds).to(device).float()
            similarities = torch.matmul(audio_embeds, text_embeddings.T.to(device))
            loss = criterion(similarities, true_labels.to(device))
            preds = similarities.argmax(dim=-1).cpu()
            correct += (preds == true_labels).sum().item()
            total += x.size(0)
            total_loss += loss.item() * x.size(0)
    return total_loss / total, correct / total


def train_residual(clap, train_loader, val_loader, text_embeds, pca_files, layers, lr=0.01, epochs=10, device=None):
    """Body of train_with_config (src/training.py:100-135) without the W&B calls. Returns (residuals, history)."""
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
