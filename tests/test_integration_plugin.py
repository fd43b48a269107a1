"""The drop-in boundary exercised for real: the ``distar/agent/b200`` pipeline (distar_b200/plugin) is found by the
reference's OWN import_helper, instantiated under the reference's OWN BaseLearner (hooks, logger, checkpoint helper, lr
scheduler unmodified) and run for a few iterations on CPU with the kernel stand-ins; the checkpoint the framework's SaveCkptHook
writes is then loaded back by its LoadCkptHook into a second learner.  Runs only where /root/reference exists."""
import os

import pytest
import torch

import ref_import

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_import.reference_available(), reason='reference tree not mounted')]


@pytest.fixture()
def pipeline(tmp_path, monkeypatch):
    ref_import.install_shims()
    from distar_b200 import ops, plugin
    plugin.register('b200')
    monkeypatch.chdir(tmp_path)                       # the framework writes ./experiments/<name>/...
    ops.enable_host_logic_testing(True)
    yield
    ops.enable_host_logic_testing(False)


def _cfg(name, load_path=''):
    from distar.ctools.utils import read_config
    cfg = read_config(os.path.join(ref_import.REFERENCE_ROOT, 'distar/bin/rl_user_config.yaml'))
    cfg.common.experiment_name = name
    cfg.common.type = 'rl'
    cfg.learner.agent = 'b200'
    cfg.learner.job_type = 'eval'                     # no league / coordinator in a unit test: synthetic batches
    cfg.learner.use_cuda = False
    cfg.learner.use_distributed = False
    assert cfg.learner.use_value_feature              # the shipped self-play default: ValueEncoder in front of the baselines
    cfg.learner.player_id = 'MP0'
    cfg.learner.load_path = load_path
    cfg.learner.load_optimizer = True                 # (bin/rl_user_config.yaml ships it off; the hook default is on)
    cfg.learner.data.batch_size = 1
    cfg.actor.traj_len = 2
    cfg.model.spatial_x = cfg.model.spatial_y = 128
    cfg.model.enable_baselines = ['winloss']
    return cfg


def test_b200_pipeline_runs_under_the_reference_learner_framework(pipeline):
    from distar.agent.import_helper import import_module
    from distar.ctools.worker.learner.base_learner import BaseLearner
    RLLearner = import_module('b200', 'RLLearner')            # how rl_train.py:43 picks the class
    assert issubclass(RLLearner, BaseLearner)
    assert import_module('b200', 'Agent').__name__ == 'Agent' and import_module('b200', 'SLLearner').__name__ == 'SLLearner'
    # the b200 Agent computes with our Model; the default pipeline's Agent, importable side by side, keeps the reference's
    import distar.agent.default.agent as default_agent
    from distar_b200.model import Model as B200Model
    assert import_module('b200', 'Agent').__init__.__globals__['Model'] is B200Model
    assert default_agent.Model is not B200Model and default_agent.Agent.__init__.__globals__['Model'] is default_agent.Model
    learner = RLLearner(_cfg('plug_a'))
    from distar_b200.model import Model
    from distar_b200.ops import FlatAdam
    assert isinstance(learner.model, Model) and isinstance(learner.optimizer, FlatAdam)
    assert isinstance(learner.lr_scheduler, torch.optim.lr_scheduler.MultiStepLR)
    # save_grad (rl_learner.py:35-47,118-130; set up by the stock constructor in train mode): per-parameter norms before / after
    # the clip go to the three tensorboard writers
    class _Writer:
        def __init__(self):
            self.rows = {}

        def add_scalar(self, k, v, global_step=None):
            self.rows[k] = v
    learner._save_grad, learner.save_log_freq = True, 1
    learner.grad_tb_logger, learner.clip_grad_tb_logger, learner.model_tb_logger = _Writer(), _Writer(), _Writer()
    w0 = learner.model.flat_param.clone()
    learner.run(max_iterations=2)                             # before_run / after_iter / after_run hooks of the framework
    names = [n for n, p in learner.model.named_parameters() if p.requires_grad]
    assert set(learner.grad_tb_logger.rows) == set(names) == set(learner.clip_grad_tb_logger.rows)
    total = sum(v * v for v in learner.grad_tb_logger.rows.values()) ** 0.5
    clipped = sum(v * v for v in learner.clip_grad_tb_logger.rows.values()) ** 0.5
    assert clipped <= total + 1e-6 and clipped <= learner.optimizer.max_norm * 1.001
    assert learner.last_iter.val == 2 and not torch.equal(w0, learner.model.flat_param)
    ckpt_dir = os.path.join('experiments', 'plug_a', 'MP0', 'checkpoint')
    files = sorted(os.listdir(ckpt_dir))
    assert files, 'SaveCkptHook (after_run) wrote nothing'
    ck = torch.load(os.path.join(ckpt_dir, files[-1]), map_location='cpu', weights_only=False)
    assert set(ck.keys()) >= {'model', 'optimizer', 'last_iter'} and ck['last_iter'] == 2
    assert set(ck['optimizer'].keys()) == {'state', 'param_groups'}
    # resume through the framework's LoadCkptHook (learner_hook.py:136-166: model + optimizer + last_iter)
    resumed = RLLearner(_cfg('plug_b', load_path=os.path.abspath(os.path.join(ckpt_dir, files[-1]))))
    resumed.call_hook('before_run')
    assert resumed.last_iter.val == 2 and resumed.optimizer.t == learner.optimizer.t == 2
    assert torch.equal(resumed.model.flat_param, learner.model.flat_param)
    assert torch.equal(resumed.optimizer.exp_avg_sq, learner.optimizer.exp_avg_sq)
    # league value reset (rl_learner.py:225-242): critic re-initialised, policy untouched, optimiser rebuilt
    before = {k: v.clone() for k, v in resumed.model.state_dict().items()}
    resumed.reset_value()
    after = resumed.model.state_dict()
    assert all(torch.equal(before[k], after[k]) for k in before if 'value' not in k)
    assert any(not torch.equal(before[k], after[k]) for k in before if k.startswith('value_networks.'))
    assert any(not torch.equal(before[k], after[k]) for k in before if k.startswith('value_encoder.'))
    assert resumed.optimizer.t == 0
    # the reference's torch.optim.Adam accepts the optimizer entry of OUR checkpoint (same per-parameter layout)
    ref_model, _cfg_ref, _mods = ref_import.load_reference(spatial=128, enable_baselines=('winloss',), use_value_feature=True)
    ref_model.load_state_dict(ck['model'], strict=True)
    ref_opt = torch.optim.Adam(ref_model.parameters(), lr=1e-5, betas=(0.0, 0.99), eps=1e-5)
    ref_opt.load_state_dict(ck['optimizer'])
    assert len(ref_opt.state_dict()['state']) == len(ck['optimizer']['state'])


def test_b200_sl_pipeline_runs_under_the_reference_learner_framework(pipeline):
    """SLLearner of the b200 pipeline under the reference's BaseLearner on the stock bin/sl_user_config.yaml (momentum_norm clip,
    warm-up scheduler, weight decay), synthetic batches (job_type != 'train'): the first 6 iterations only carry the LSTM state
    (sl_learner.py:64-72), then the weights move; the checkpoint hook's file loads back."""
    from distar.agent.import_helper import import_module
    from distar.ctools.utils import read_config
    from distar.ctools.worker.learner.base_learner import BaseLearner
    cfg = read_config(os.path.join(ref_import.REFERENCE_ROOT, 'distar/bin/sl_user_config.yaml'))
    cfg.common.experiment_name = 'plug_sl'
    cfg.common.type = 'sl'
    cfg.learner.agent = 'b200'
    cfg.learner.job_type = 'eval'
    cfg.learner.use_cuda = False
    cfg.learner.data.batch_size = 1
    cfg.learner.data.trajectory_length = 2
    cfg.model = {'spatial_x': 128, 'spatial_y': 128}
    SLLearner = import_module('b200', 'SLLearner')
    assert issubclass(SLLearner, BaseLearner)
    learner = SLLearner(cfg)
    from distar_b200.ops import FlatAdam
    assert isinstance(learner.optimizer, FlatAdam) and learner.optimizer.clip_type == 'momentum_norm'
    w0 = learner.model.flat_param.clone()
    learner.run(max_iterations=6)
    assert torch.equal(w0, learner.model.flat_param)              # iterations 0..5: no update yet
    learner.run(max_iterations=2)
    assert learner.last_iter.val == 8 and not torch.equal(w0, learner.model.flat_param)
    # warm-up of the stock config (use_warmup, 20000 steps).  The scheduler is stepped by the framework's LrSchdulerHook after
    # every iteration (learner_hook.py:111-112) AND by _train after every optimiser step (sl_learner.py:70), as in the reference:
    # 8 + 2 steps so far
    from distar.ctools.torch_utils.lr_scheduler_util import GradualWarmupScheduler
    assert isinstance(learner.lr_scheduler, GradualWarmupScheduler)
    assert abs(learner.optimizer.param_groups[0]['lr'] - 1e-3 * 10 / 20000) < 1e-12
    ckpt_dir = os.path.join('experiments', 'plug_sl', 'checkpoint')
    found = [os.path.join(r, f) for r, _d, fs in os.walk(os.path.join('experiments', 'plug_sl')) for f in fs if f.endswith('.pth.tar') or f.endswith('.pth')]
    assert found, 'SaveCkptHook wrote nothing (looked under %s)' % ckpt_dir
    ck = torch.load(found[-1], map_location='cpu', weights_only=False)
    assert set(ck.keys()) >= {'model', 'optimizer', 'last_iter'}


def test_stock_agent_runs_on_the_b200_model(pipeline):
    """The reference's Agent (agent.py:97-145), built through the b200 pipeline on the stock config at its default 160 x 152 map
    size: the constructor's realtime warm-up, then compute_logp_action / compute_teacher_logit on the reference's own
    fake_step_data return exactly the structure (keys, shapes, dtypes) the reference model returns for the same input."""
    from distar.agent.import_helper import import_module
    from distar.agent.default.lib.features import fake_step_data
    from distar.ctools.utils import read_config
    import distar.agent.default.agent as default_agent
    cfg = read_config(os.path.join(ref_import.REFERENCE_ROOT, 'distar/bin/rl_user_config.yaml'))
    cfg.common.type = 'rl'                         # as bin/rl_train.py sets it
    cfg.actor.use_cuda = False
    cfg.actor.job_type = 'train'                   # -> a teacher model too (agent.py:142-143)
    cfg.env.realtime = True                        # -> warm-up forward inside the constructor (agent.py:120-127)
    cfg.agent.z_path = 'unused.json'
    from distar.agent.default.lib import features as F
    F.SPATIAL_SIZE[:] = [152, 160]
    agent = import_module('b200', 'Agent')(cfg)
    from distar_b200.model import Model
    assert isinstance(agent.model, Model) and isinstance(agent.teacher_model, Model)
    assert (agent.model.spatial_x, agent.model.spatial_y) == (160, 152) and agent._hidden_size == 384 and agent._num_layers == 3
    ref_model = default_agent.Model(cfg)

    def structure(tree, path=''):
        if torch.is_tensor(tree):
            return {path: (tuple(tree.shape), tree.dtype)}
        out = {}
        items = tree.items() if isinstance(tree, dict) else enumerate(tree)
        for k, v in items:
            out.update(structure(v, '%s/%s' % (path, k)))
        return out
    obs = fake_step_data(share_memory=True, batch_size=2, hidden_size=384, hidden_layer=3, train=False)
    with torch.no_grad():
        torch.manual_seed(0)
        mine = agent.model.compute_logp_action(**obs)
        torch.manual_seed(0)
        want = ref_model.compute_logp_action(**obs)
    sm, sw = structure(mine), structure(want)
    # the number of pointer-network steps depends on the sampled selection (different weights): compare with that dim freed
    free = lambda d: {k: ((s[0],) + s[2:] if k == '/logit/selected_units' else s, t) for k, (s, t) in d.items()}
    assert free(sm) == free(sw)
    tobs = fake_step_data(share_memory=True, batch_size=2, hidden_size=384, hidden_layer=3, train=True)
    with torch.no_grad():
        mine_t = agent.teacher_model.compute_teacher_logit(**tobs)
        want_t = ref_model.compute_teacher_logit(**tobs)
    assert structure(mine_t) == structure(want_t)


def test_pipeline_exposes_every_class_import_helper_can_ask_for(pipeline):
    """import_helper.MODULE_PATHS: RLLearner, SLLearner, Agent, ReplayDecoder.  The replay decoder is game-protocol CPU code the
    pipeline re-exports from the reference; this image lacks its third-party parsers (mpyq ...), so the lookup must get as far
    as the STOCK module's own imports."""
    from distar.agent.import_helper import MODULE_PATHS, import_module
    assert set(MODULE_PATHS) == {'RLLearner', 'SLLearner', 'Agent', 'ReplayDecoder'}
    for name in ('RLLearner', 'SLLearner', 'Agent'):
        assert import_module('b200', name).__name__ == name
    try:
        cls = import_module('b200', 'ReplayDecoder')
    except ModuleNotFoundError as e:
        assert not (e.name or '').startswith('distar.agent.b200') and not (e.name or '').startswith('distar_b200'), e
    else:
        import distar.agent.default.replay_decoder as stock
        assert cls is stock.ReplayDecoder
