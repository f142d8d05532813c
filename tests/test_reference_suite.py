"""The reference's own unit tests, run against the drop-in classes (same names, constructor arguments and
assertions; only the import root differs):
  tests/rollout_buffer_test.py:12-118   CoupledTensor indexing known answers, gradient flow through the buffer
  tests/linearisation_test.py:8-35      one real step lands inside Box(linearised next state, E G_noise)
  tests/safeguard_test.py:10-37         BP / ray-mask `.actions()` are differentiable w.r.t. the action
"""
import pytest
import torch

from safegbpo_b200.learning_algorithms.components.coupled_buffer import CoupledBuffer
from safegbpo_b200.learning_algorithms.components.coupled_tensor import CoupledTensor


def init_coupled_tensor():
    ct = CoupledTensor(3, 6, 2)
    ct.tensor = torch.arange(3 * 6 * 2).reshape(3, 6, 2).to(torch.get_default_dtype())
    return ct


def test_coupled_tensor_coupled_indexing():
    ct = init_coupled_tensor()
    t = torch.ones(6, dtype=torch.int64)
    t[2] = 2
    retrieval_coupled = ct[t]
    expected_result = torch.tensor([[12., 13.], [14., 15.], [28., 29.], [18., 19.], [20., 21.], [22., 23.]])
    assert retrieval_coupled.shape == torch.Size([6, 2])
    assert torch.allclose(retrieval_coupled, expected_result)
    ct[t] = torch.ones(6, 2)
    assert torch.allclose(ct[t], torch.ones(6, 2))


def test_coupled_tensor_full_indexing():
    ct = init_coupled_tensor()
    assert ct[1, 4, 1] == 21.0
    ct[2, 5, 1] = 100
    assert ct[2, 5, 1] == 100.0


def test_coupled_tensor_coupled_boolean_masking():
    ct = init_coupled_tensor()
    t = torch.ones(6, dtype=torch.int64)
    t[2] = 2
    mask = torch.tensor([True, False, True, False, False, False])
    retrieval_coupled = ct[t[mask], mask]
    assert retrieval_coupled.shape == torch.Size([2, 2])
    assert torch.allclose(retrieval_coupled, torch.tensor([[12., 13.], [28., 29.]]))
    ct[t[mask], mask] = torch.ones(2, 2)
    assert torch.allclose(ct[t[mask], mask], torch.ones(2, 2))


def test_coupled_tensor_nonzero_count_and_calc():
    ct = init_coupled_tensor()
    assert torch.allclose(ct.nonzero(), ct.tensor.nonzero())
    assert ct.count_nonzero() == ct.tensor.count_nonzero()
    assert torch.allclose(ct * 2, ct.tensor * 2)
    assert torch.allclose(ct + 2, ct.tensor + 2)
    assert torch.allclose(ct - 2, ct.tensor - 2)


def test_coupled_tensor_gradients():
    ct = init_coupled_tensor()
    t = torch.ones(6, dtype=torch.int64)
    gradient_tacker = torch.ones((6, 2), requires_grad=True)
    ct[t] = gradient_tacker
    loss = torch.zeros(6)
    loss -= (ct[t] * 2).sum(dim=1)
    loss.backward(torch.ones([6]))
    assert gradient_tacker.grad[0, 0] == -2.0


def test_coupled_buffer_gradients():
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        cb = CoupledBuffer(10, 6, 2, True)
        gradient_tacker_obs = torch.ones((6, 2), requires_grad=True)
        gradient_tacker_rew = torch.ones(6, requires_grad=True)
        target_values = torch.ones(6, requires_grad=False)
        episode_ends = torch.zeros(6, dtype=torch.bool, requires_grad=False)
        dummy_obs, dummy_rew = torch.zeros(6, 2), torch.zeros(6)
        dummy_target, dummy_episode_ends = torch.zeros(6), torch.zeros(6, dtype=torch.bool)
        cb.add(dummy_obs, dummy_rew, dummy_episode_ends, dummy_target)
        cb.add(gradient_tacker_obs, gradient_tacker_rew, episode_ends, target_values)
        cb.add(dummy_obs, dummy_rew, dummy_episode_ends, dummy_target)
        loss = -(cb.observations * 2).sum(dim=[1, 2]) - cb.rewards.sum(dim=1)
        loss.backward(torch.ones_like(loss))
        assert gradient_tacker_obs.grad[0, 0] == -2.0
        assert gradient_tacker_rew.grad[0] == -1.0
        cb.reset()
        gradient_tacker_obs = torch.ones((6, 2), requires_grad=True)
        gradient_tacker_rew = torch.ones(6, requires_grad=True)
        cb.add(gradient_tacker_obs, gradient_tacker_rew, episode_ends, target_values)
        loss = -(cb.observations * 2).sum(dim=[1, 2]) - cb.rewards.sum(dim=1)
        loss.backward(torch.ones_like(loss))
        assert gradient_tacker_obs.grad[0, 0] == -2.0
        assert gradient_tacker_rew.grad[0] == -1.0
    finally:
        torch.set_default_dtype(prev)


@pytest.fixture
def cuda_f64_default():
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    torch.set_default_device("cuda:0")
    yield torch.device("cuda:0")
    torch.set_default_device("cpu")
    torch.set_default_dtype(prev)


def _linearisation_test(env_class):
    from safegbpo_b200.sets.box import Box
    env = env_class(num_envs=2, num_steps=1000)
    env.reset()
    constant_mat, state_mat, action_mat, noise_mat = env.linear_dynamics()
    action = torch.ones(env.num_envs, env.action_dim) * 0.1
    lin_next_state = (constant_mat + torch.einsum('bij,bj->bi', state_mat, env.state - env.state)
                      + torch.einsum('bij,bj->bi', action_mat, action)
                      + torch.einsum('bij,bj->bi', noise_mat, env.noise_set.center - env.noise_set.center))
    lin_next_state_generator = torch.matmul(noise_mat, env.noise_set.generator)
    lin_next_state = torch.clamp(lin_next_state, env.state_set.min, env.state_set.max)
    next_state_box = Box(lin_next_state, lin_next_state_generator)
    env.step(action)
    next_state = env.state
    assert next_state_box.contains(next_state).all() or torch.allclose(lin_next_state, next_state, rtol=0.01)


@pytest.mark.gpu
def test_pendulum_linearisation(cuda_f64_default):
    from safegbpo_b200.envs.simulators.pendulum import PendulumEnv
    _linearisation_test(PendulumEnv)


@pytest.mark.gpu
def test_household_linearisation(cuda_f64_default):
    from safegbpo_b200.envs.simulators.household import HouseholdEnv
    _linearisation_test(HouseholdEnv)


@pytest.mark.gpu
def test_boundary_projection(cuda_f64_default):
    from safegbpo_b200.envs.balance_pendulum import BalancePendulumEnv
    from safegbpo_b200.safeguards.boundary_projection import BoundaryProjectionSafeguard
    env = BalancePendulumEnv(num_envs=2, num_steps=100)
    env.reset()
    wrapper = BoundaryProjectionSafeguard(env)
    action = torch.tensor([[0.0], [1.0]], requires_grad=True)
    safe_actions = wrapper.actions(action)
    safe_actions.backward(torch.ones_like(safe_actions))
    assert torch.count_nonzero(action.grad) == 2


@pytest.mark.gpu
def test_ray_mask(cuda_f64_default):
    from safegbpo_b200.envs.balance_pendulum import BalancePendulumEnv
    from safegbpo_b200.safeguards.ray_mask import RayMaskSafeguard
    env = BalancePendulumEnv(num_envs=2, num_steps=100)
    env.reset()
    for zonotopic_approximation in [True, False]:
        wrapper = RayMaskSafeguard(env, linear_projection=True, zonotopic_approximation=zonotopic_approximation,
                                   passthrough=False)
        action = torch.tensor([[0.0], [1.0]], requires_grad=True)
        safe_actions = wrapper.actions(action)
        safe_actions.backward(torch.ones_like(safe_actions))
        assert torch.count_nonzero(action.grad) == 2
