"""TEST INFRASTRUCTURE ONLY - loads the *reference's own* functions for the hot path.

Runs only in the build container (needs /root/reference, which does not exist on
the GPU box).  It is used by ``oracle/make_golden.py`` to produce the committed
fixtures under ``tests/golden/`` and by ``tests/test_oracle_vs_reference.py`` (skipped
when the reference tree is absent).  Nothing here is copied into the product:
functions are pulled out of the reference files *by line range at run time* into a
temp module, because the enclosing modules import the absent Isaac Gym bindings
(SURVEY.md section 8c).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
import textwrap
import types

REF = os.environ.get("EMLOCO_REFERENCE", "/root/reference")
PACER = os.path.join(REF, "pacer", "pacer")


def available() -> bool:
    return os.path.isdir(PACER)


def _lines(path, a, b):
    with open(path) as f:
        L = f.readlines()
    return "".join(L[a - 1:b])


_cache = {}


def _import_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Returns a namespace with the reference functions (torch, CPU)."""
    if "ns" in _cache:
        return _cache["ns"]
    import numpy as np
    if not hasattr(np, "float"):
        np.float = float  # isaacgym/torch_utils.py:135 default arg
    if not hasattr(np, "int"):
        np.int = int
    import torch

    # stub packages so that `from isaacgym.torch_utils import *` resolves
    isaac = types.ModuleType("isaacgym")
    isaac.__path__ = []
    sys.modules.setdefault("isaacgym", isaac)
    itu = _import_path("isaacgym.torch_utils",
                       os.path.join(REF, "isaacgym/python/isaacgym/torch_utils.py"))
    isaac.torch_utils = itu
    utils_pkg = types.ModuleType("utils")
    utils_pkg.__path__ = []
    sys.modules.setdefault("utils", utils_pkg)
    ptu = _import_path("utils.torch_utils", os.path.join(PACER, "utils/torch_utils.py"))
    utils_pkg.torch_utils = ptu

    tmp = tempfile.mkdtemp(prefix="emloco_ref_")
    hum = os.path.join(PACER, "env/tasks/humanoid.py")
    amp = os.path.join(PACER, "env/tasks/humanoid_amp.py")
    ter = os.path.join(PACER, "env/tasks/humanoid_pedestrain_terrain.py")
    header = ("import torch\nimport numpy as np\nfrom isaacgym.torch_utils import *\n"
              "from utils import torch_utils\nfrom typing import Tuple, List\n\n")
    # remove_base_rot is referenced (dead branch, upright=True) by the jit functions
    body = header
    body += ("@torch.jit.script\ndef remove_base_rot(quat):\n"
             "    base_rot = quat_conjugate(torch.tensor([[0.5, 0.5, 0.5, 0.5]]).to(quat))\n"
             "    shape = quat.shape[0]\n"
             "    return quat_mul(quat, base_rot.repeat(shape, 1))\n\n")
    body += _lines(hum, 1327, 1338)       # dof_to_obs_smpl
    body += "\n" + _lines(hum, 1626, 1687)  # compute_humanoid_observations_smpl_max
    body += "\n" + _lines(amp, 917, 971)    # build_amp_observations_smpl
    body += "\n" + _lines(ter, 1468, 1530)  # compute_humanoid_reset
    body += "\n" + _lines(ter, 1533, 1538)  # quat_apply_yaw
    body += "\n" + _lines(ter, 1549, 1592)  # compute_location_observations / reward
    p = os.path.join(tmp, "emloco_ref_jit.py")
    with open(p, "w") as f:
        f.write(body)
    jit = _import_path("emloco_ref_jit", p)

    # plain-python methods, re-hosted on small holder classes
    tg = os.path.join(PACER, "env/util/traj_generator.py")
    ca = os.path.join(PACER, "learning/common_agent.py")
    ac = os.path.join(PACER, "learning/amp_continuous.py")
    meth = header + "class TrajHolder:\n" + _lines(tg, 278, 296)
    meth += "\nclass TerrainHolder:\n" + _lines(ter, 1212, 1218) + "\n" + _lines(ter, 1221, 1226)
    # non-group branch of sample_height_points (ter:1282-1288)
    meth += ("        heights1 = heightsamples[px, py]\n"
             "        heights2 = heightsamples[px + 1, py + 1]\n"
             "        heights = torch.min(heights1, heights2)\n"
             "        return heights * self.vertical_scale\n")
    meth += "\nclass AgentHolder:\n" + _lines(ca, 573, 587)
    p2 = os.path.join(tmp, "emloco_ref_meth.py")
    with open(p2, "w") as f:
        f.write(meth)
    methm = _import_path("emloco_ref_meth", p2)

    # TrajGenerator.reset (traj_generator.py:60-237) on a holder class; its module-level `torch` / `random` names are
    # swapped by make_golden for proxies that replay recorded uniform draws
    p3 = os.path.join(tmp, "emloco_ref_trajreset.py")
    with open(p3, "w") as f:
        f.write("import numpy as np\nimport random\nimport torch\n\nclass TrajResetHolder:\n" + _lines(tg, 60, 237)
                + "\n" + _lines(tg, 261, 262))
    trajreset = _import_path("emloco_ref_trajreset", p3)

    # ValuePoseNet with a stub matplotlib
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.__path__ = []
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    vpn = _import_path("emloco_ref_vpn", os.path.join(PACER, "learning/value_pose_net.py"))
    rms = _import_path("emloco_ref_rms", os.path.join(PACER, "utils/running_mean_std.py"))

    ns = types.SimpleNamespace(
        torch=torch, itu=itu, ptu=ptu, jit=jit, meth=methm, trajreset=trajreset,
        ValuePoseNet=vpn.ValuePoseNet, RunningMeanStd=rms.RunningMeanStd,
        left_to_right_index=[0, 5, 6, 7, 8, 1, 2, 3, 4, 9, 10, 11, 12, 13, 19, 20, 21, 22, 23,
                             14, 15, 16, 17, 18],  # humanoid.py:334
    )
    _cache["ns"] = ns
    return ns


# ------------------------------------------------------------------------------------------------------------------
# The reference's own network classes (learning/network_builder.py -> amp_network_builder.py -> amp_network_sept_builder.py
# -> amp_network_sept_value_builder.py).  They import rl_games (rl-games==1.1.4, pinned in pacer/requirements.txt, NOT vendored
# and not installed): the five names they touch at import / build time are stubbed below - an object factory (a dict of
# builders) and four symbols that are only referenced by branches the default cfg never takes (noisy dense, D2RL, SAC, conv).
# The network arithmetic that runs is the reference's.
# ------------------------------------------------------------------------------------------------------------------
def _stub_rl_games():
    if "rl_games" in sys.modules and getattr(sys.modules["rl_games"], "_emloco_stub", False):
        return
    import torch

    def pkg(name):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        return m

    rl = pkg("rl_games"); rl._emloco_stub = True
    common = pkg("rl_games.common"); algos = pkg("rl_games.algos_torch")
    rl.common, rl.algos_torch = common, algos

    of = types.ModuleType("rl_games.common.object_factory")

    class ObjectFactory:                       # rl_games/common/object_factory.py (1.1.4): name -> builder(**kwargs)
        def __init__(self):
            self._builders = {}

        def register_builder(self, name, builder):
            self._builders[name] = builder

        def set_builders(self, builders):
            self._builders = builders

        def create(self, name, **kwargs):
            builder = self._builders.get(name)
            if not builder:
                raise ValueError(name)
            return builder(**kwargs)

    of.ObjectFactory = ObjectFactory
    sys.modules["rl_games.common.object_factory"] = of
    common.object_factory = of
    for sub, names in (("torch_ext", ()), ("layers", ()), ("d2rl", ("D2RLNet",)), ("sac_helper", ("SquashedNormal",))):
        m = types.ModuleType(f"rl_games.algos_torch.{sub}")
        for n in names:
            setattr(m, n, type(n, (), {}))
        sys.modules[f"rl_games.algos_torch.{sub}"] = m
        setattr(algos, sub, m)


def network_params():
    """The `params.network` tree of data/cfg/train/rlg/amp_humanoid_smpl_sept_task.yaml (read from the reference at run time)."""
    import yaml
    with open(os.path.join(PACER, "data/cfg/train/rlg/amp_humanoid_smpl_sept_task.yaml")) as f:
        return yaml.safe_load(f)["params"]["network"]


def load_network(seed=0, randomize_bias=True):
    """AMPSeptValueBuilder.Network built exactly as `AMPAgent._build_net_config` + `ModelAMPContinuous.build` do for the default
    task (amp_continuous.py:500-510, common_agent.py:56-63,611-619): obs 1422 = 368 + 1054 (traj 30 + heightmap 1024), AMP obs 3090,
    69 actions.  randomize_bias: the reference zeroes every bias at construction; the fixtures use non-zero ones so that the
    tests are not blind to them."""
    if ("net", seed, randomize_bias) in _cache:
        return _cache[("net", seed, randomize_bias)]
    R = load()
    torch = R.torch
    _stub_rl_games()
    learning = types.ModuleType("learning")
    learning.__path__ = [os.path.join(PACER, "learning")]
    sys.modules["learning"] = learning
    import importlib
    importlib.import_module("learning.network_builder")
    mod = importlib.import_module("learning.amp_network_sept_value_builder")
    import contextlib, io
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        mean_std = R.RunningMeanStd((1422,))
        builder = mod.AMPSeptValueBuilder()
        builder.load(network_params())
        net = builder.build("amp", actions_num=69, input_shape=(1422,), num_seqs=1, value_size=1, amp_input_shape=(3090,),
                            self_obs_size=368, task_obs_size=1054, task_obs_size_detail={"traj": 30, "heightmap": 1024},
                            mean_std=mean_std)
    if randomize_bias:
        with torch.no_grad():
            for m in net.modules():
                if isinstance(m, torch.nn.Linear):
                    m.bias.uniform_(-0.1, 0.1)
    net.eval()
    _cache[("net", seed, randomize_bias)] = net
    return net


def load_agent_blocks():
    """Pieces of the agent classes re-hosted on a holder (the enclosing modules import rl_games / isaacgym / tensorboardX):
      * AMPAgent._preproc_obs (learning/amp_continuous.py:322-333),
      * AMPAgent._preproc_amp_obs / _combine_rewards / _eval_disc / _calc_amp_rewards / _calc_disc_rewards
        (learning/amp_continuous.py:651-693),
      * CommonAgent._eval_critic, _actor_loss, _critic_loss, _calc_advs, bound_loss (learning/common_agent.py:594-602,647-696),
      * AMPAgent._disc_loss .. _compute_disc_acc (learning/amp_continuous.py:536-616), _sym_loss (:517-534),
      * AMPValueAgent._task_value_loss (learning/amp_continuous_value.py:430-444),
      * AMPValueAgent.calc_gradients up to the total loss (learning/amp_continuous_value.py:277-362) as `Holder.calc_loss`,
      * ModelAMPContinuous.Network.dropout_amp_obs / get_dropout_mask (learning/amp_models.py:46-90),
      * AMPValueAgent.calc_gradients up to the total loss (learning/amp_continuous_value.py:277-362) as `Holder.calc_loss`,
      * ModelAMPContinuous.Network.dropout_amp_obs / get_dropout_mask (learning/amp_models.py:46-90),
      * the per-step body of AMPValueAgent.play_steps from env_step to the end of the no_grad block
        (learning/amp_continuous_value.py:61-121) as `Holder.play_block(self)`."""
    if "agent" in _cache:
        return _cache["agent"]
    load()
    ac = os.path.join(PACER, "learning/amp_continuous.py")
    ca = os.path.join(PACER, "learning/common_agent.py")
    av = os.path.join(PACER, "learning/amp_continuous_value.py")
    src = "import torch\nimport numpy as np\nfrom torch import nn\n\nclass Holder:\n"
    src += _lines(ac, 322, 333) + "\n" + _lines(ac, 517, 616) + "\n" + _lines(ac, 651, 693) + "\n"
    src += _lines(ca, 594, 602) + "\n" + _lines(ca, 647, 696) + "\n"
    src += _lines(av, 430, 444) + "\n"
    # locals of play_steps that the block reads (n, res_dict, terminated_flags, reward_raw) become arguments; its locals are returned
    # calc_gradients from set_train() to the assembled total loss (learning/amp_continuous_value.py:277-362): locals returned
    src += ("    def calc_loss(self, input_dict):\n" + textwrap.indent(textwrap.dedent(_lines(av, 277, 362)), " " * 8)
            + "\n        return locals()\n")
    # ModelAMPContinuous.Network.dropout_amp_obs / get_dropout_mask (learning/amp_models.py:46-90)
    am = os.path.join(PACER, "learning/amp_models.py")
    src += textwrap.indent(textwrap.dedent(_lines(am, 46, 90)), " " * 4) + "\n"
    src += ("    def play_block(self, n, res_dict, terminated_flags, reward_raw):\n"
            + textwrap.indent(textwrap.dedent(_lines(av, 61, 121)), " " * 8) + "\n        return locals()\n")
    tmp = tempfile.mkdtemp(prefix="emloco_ref_")
    p = os.path.join(tmp, "emloco_ref_agent.py")
    with open(p, "w") as f:
        f.write(src)
    mod = _import_path("emloco_ref_agent", p)
    _cache["agent"] = mod.Holder
    return mod.Holder


PLAYER_LOCALS = ("action", "real_traj", "inverted_envs", "cr", "steps", "c_task_value", "c_critic_value", "c_disc_reward", "c_loc_reward",
                 "c_pow_reward", "rew_disc_coef", "max_frame_rew", "min_frame_rew", "rew_lists", "bar", "t", "games_played", "cr_to_pred",
                 "rewards_loc", "rewards_pow", "rewards_disc", "min_reward", "max_reward", "total_value_loss", "valuenet_pred", "waypoint_traj",
                 "init_pose", "init_vel")


def load_player_block():
    """One iteration of the step loop of AMPPlayerContinuousValue.run (learning/amp_value_players.py:123-198: env_step, inversion
    penalty, LocoVal score at n == 0, discounted reward accumulation, snapshot at step_to_pred / early done, MSE against the
    normalised return) as `Holder.player_block(self, n, L)`; L carries the loop's local variables (PLAYER_LOCALS)."""
    if "player" in _cache:
        return _cache["player"]
    load()
    pl = os.path.join(PACER, "learning/amp_value_players.py")
    src = "import copy\nimport torch\nimport numpy as np\n\nclass Holder:\n    def player_block(self, n, L):\n"
    src += "".join(f"        {v} = L.get('{v}')\n" for v in PLAYER_LOCALS)
    src += "        done_count = 0\n"
    src += textwrap.indent(textwrap.dedent(_lines(pl, 123, 198)), " " * 8) + "\n"
    src += "        L.update({" + ", ".join(f"'{v}': {v}" for v in PLAYER_LOCALS) + ", 'done': done, 'done_count': done_count})\n"
    src += "        if done_count > 0:\n            L.update(norm_rewards=norm_rewards, value_loss=value_loss)\n        return L\n"
    tmp = tempfile.mkdtemp(prefix="emloco_ref_")
    p = os.path.join(tmp, "emloco_ref_player.py")
    with open(p, "w") as f:
        f.write(src)
    mod = _import_path("emloco_ref_player", p)
    _cache["player"] = mod.Holder
    return mod.Holder


def load_motion_lib_block():
    """MotionLibSMPL.get_motion_state_smpl / _calc_frame_blend / _get_num_bodies / _local_rotation_to_dof_smpl
    (utils/motion_lib_smpl.py:485-563,596-614) on a holder (the module imports poselib, joblib pickles and the SMPL parser)."""
    if "motionlib" in _cache:
        return _cache["motionlib"]
    load()
    ml = os.path.join(PACER, "utils/motion_lib_smpl.py")
    src = "import torch\nimport numpy as np\nfrom utils import torch_utils\n\nclass Holder:\n" + _lines(ml, 485, 563) + "\n" + _lines(ml, 596, 614) + "\n"
    tmp = tempfile.mkdtemp(prefix="emloco_ref_")
    p = os.path.join(tmp, "emloco_ref_motionlib.py")
    with open(p, "w") as f:
        f.write(src)
    mod = _import_path("emloco_ref_motionlib", p)
    _cache["motionlib"] = mod.Holder
    return mod.Holder


def load_pd_block():
    """Humanoid._build_pd_action_offset_scale (env/tasks/humanoid.py:949-1025) as a method of a holder: the PD target offset / scale
    table the actions are mapped with (a2).  The holder carries what the method reads: _dof_offsets, dof_limits_lower / _upper
    (Isaac Gym's DOF properties = the MJCF joint ranges), _bias_offset, smpl_humanoid, _dof_names, _has_smpl_pd_offset,
    _has_upright_start, device."""
    if "pd" in _cache:
        return _cache["pd"]
    R = load()
    hp = os.path.join(PACER, "env/tasks/humanoid.py")
    src = "import numpy as np\nimport torch\nfrom isaacgym.torch_utils import to_torch\n\nclass Holder:\n" + _lines(hp, 949, 1025) + "\n"
    tmp = tempfile.mkdtemp(prefix="emloco_ref_")
    p = os.path.join(tmp, "emloco_ref_pd.py")
    with open(p, "w") as f:
        f.write(src)
    mod = _import_path("emloco_ref_pd", p)
    _cache["pd"] = mod.Holder
    return mod.Holder


def load_plausibl_mlp():
    """plausibl/test_value_mlp.py:24-113 `class MLP` (the script's imports point at a developer's home directory)."""
    if "plausibl" in _cache:
        return _cache["plausibl"]
    load()
    src = "import torch\nimport torch.nn as nn\n\n" + _lines(os.path.join(REF, "plausibl/test_value_mlp.py"), 24, 113)
    tmp = tempfile.mkdtemp(prefix="emloco_ref_")
    p = os.path.join(tmp, "emloco_ref_plausibl.py")
    with open(p, "w") as f:
        f.write(src)
    mod = _import_path("emloco_ref_plausibl", p)
    _cache["plausibl"] = mod.MLP
    return mod.MLP
