"""Default hyper-parameters = the reference's argparse defaults (src/arguments.py:6-144), as a namespace."""
from types import SimpleNamespace


def default_args(**overrides):
    a = dict(
        domain_name="carla", task_name="drive", frame_stack=3, action_repeat=4, episode_length=600,
        algorithm="sgsac", discount=0.99, batch_size=128, hidden_dim=1024,
        actor_lr=1e-3, actor_beta=0.9, actor_log_std_min=-10.0, actor_log_std_max=2.0, actor_update_freq=2,
        critic_lr=1e-3, critic_beta=0.9, critic_tau=0.01, critic_target_update_freq=2, critic_weight_decay=0.0,
        num_shared_layers=11, num_head_layers=0, num_filters=32, projection_dim=100, encoder_tau=0.05,
        init_temperature=0.1, alpha_lr=1e-4, alpha_beta=0.5,
        aux_lr=3e-4, aux_beta=0.9, aux_update_freq=2,
        soda_batch_size=256, soda_tau=0.005, svea_alpha=0.5, svea_beta=0.5, sgqn_quantile=0.5, consistency=1, alpha_blending=0.2,
        seed=10081, log_dir="logs", image_size=84, image_crop_size=84,
    )
    a.update(overrides)
    if a["algorithm"] in {"rad", "curl", "pad", "soda"} and "image_size" not in overrides:
        a["image_size"] = 100                          # arguments.py:137-142
    return SimpleNamespace(**a)
