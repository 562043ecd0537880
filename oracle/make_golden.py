"""Writes tests/golden/oracle_small.npz from the oracle itself (run: python -m oracle.make_golden).
Self-generated fixture: the reference has no golden vectors and cannot run here."""
import os
import sys

import numpy as np


def golden_case(satk, root):
    from oracle import model as OR
    hp = satk.load_hparams(os.path.join(root, "examples", "ljspeech_self-attention-tacotron.json"))
    d = satk.dims_from_hparams(hp)
    ps = satk.ParamStore(d).init(2024, "random")
    f, l = satk.synthetic_batch(hp, 2, 12, 16, seed=2024)
    masks = satk.make_masks(d, 2, 12, 8, seed=2024)
    tr = OR.OracleTrainer(d, hp, ps.as_dict())
    out, grads, _ = tr.loss_and_grads(f, l, masks, True)
    return out, grads


def golden_predict_case(satk, root):
    """Free-running decode (PREDICT branch) and the transition-agent + cumulative-weights variant, frozen in oracle_predict.npz."""
    import torch
    from oracle import model as OR
    res = {}
    for tag, ov in (("base", None), ("agent_cum", "use_forward_attention_transition_agent=True,cumulative_weights=True")):
        hp = satk.load_hparams(os.path.join(root, "examples", "ljspeech_self-attention-tacotron.json"), ov)
        d = satk.dims_from_hparams(hp)
        ps = satk.ParamStore(d).init(2025, "random")
        f, l = satk.synthetic_batch(hp, 2, 11, 16, seed=2025)
        with torch.no_grad():
            pr = OR.model_predict(ps.as_dict(), d, f, max_iters=7, use_stop_token=False)
        res[tag + "_mel"] = pr["mel"].numpy()
        res[tag + "_stop"] = pr["stop"].numpy()
        res[tag + "_alignment"] = pr["alignment"].numpy()
        if tag == "agent_cum":
            masks = satk.make_masks(d, 2, 11, 8, seed=2025)
            out, grads, _ = OR.OracleTrainer(d, hp, ps.as_dict()).loss_and_grads(f, l, masks, True)
            res["agent_cum_train_loss"] = np.float64(float(out["loss"].detach()))
            res["agent_cum_grad_agent_W"] = grads["att1.agent.W"].numpy()
            res["agent_cum_grad_loc_conv_W"] = grads["att1.loc_conv.W"].numpy()
    return res


def golden_l2_case(satk, root):
    """use_l2_regularization (models/models.py:470-478) on the single-attention model (BASELINE configs[0]), frozen in oracle_l2.npz."""
    from oracle import model as OR
    hp = satk.load_hparams(os.path.join(root, "examples", "ljspeech_tacotron.json"), "use_l2_regularization=True,l2_regularization_weight=1e-4")
    d = satk.dims_from_hparams(hp)
    ps = satk.ParamStore(d).init(2026, "random")
    f, l = satk.synthetic_batch(hp, 2, 10, 12, seed=2026)
    masks = satk.make_masks(d, 2, 10, 6, seed=2026)
    out, grads, _ = OR.OracleTrainer(d, hp, ps.as_dict()).loss_and_grads(f, l, masks, True)
    return dict(loss=np.float64(float(out["loss"].detach())), regularization_loss=np.float64(float(out["regularization_loss"].detach())),
                grad_enc_prenet0_W_slice=grads["enc.prenet0.W"].numpy()[:32, :32],      # regularised
                grad_out_proj_W_slice=grads["dec.out_proj.W"].numpy()[:32, :32],        # black-listed in the ExtendedDecoder
                grad_att1_v=grads["att1.v"].numpy())


if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import satk_path
    satk = satk_path.load()
    out, grads = golden_case(satk, root)
    np.savez_compressed(os.path.join(root, "tests", "golden", "oracle_small.npz"), mel=out["mel"].detach().numpy(),
                        stop=out["stop"].detach().numpy(), alignment=out["alignment"].detach().numpy(),
                        alignment2=out["alignment2"].detach().numpy(), loss=float(out["loss"].detach()),
                        grad_dec_lstm1_W_slice=grads["dec.lstm1.W"].numpy()[100:164, :64], grad_att1_v=grads["att1.v"].numpy())
    np.savez_compressed(os.path.join(root, "tests", "golden", "oracle_predict.npz"), **golden_predict_case(satk, root))
    np.savez_compressed(os.path.join(root, "tests", "golden", "oracle_l2.npz"), **golden_l2_case(satk, root))
    print("written")
