"""Mirror of the zero-shot evaluation entry point of the reference's src/evaluation.py (evaluate_zero_shot :74-109).
The sklearn/seaborn reporting (visualize_eval_metrics :132-198) is host-side plotting and out of scope."""
import torch


def evaluate_zero_shot(model, dataloader, text_embeddings, device):
    """Returns (predictions, targets, similarities[N, classes]); inputs are int16-quantised on the device."""
    model.eval()
    all_preds, all_targets, all_similarities = [], [], []
    with torch.no_grad():
        for x, true_labels in dataloader:
            audio_embeds = model.get_audio_embedding_from_data(x=x.squeeze(1), use_tensor=False)
            audio_embeds = torch.as_tensor(audio_embeds).to(device).float()
            similarities = torch.matmul(audio_embeds, text_embeddings.T.to(device))
            all_preds.extend(similarities.argmax(dim=-1).cpu().tolist())
            all_targets.extend(true_labels.tolist())
            all_similarities.append(similarities.cpu())
    return all_preds, all_targets, torch.cat(all_similarities, dim=0).numpy()
