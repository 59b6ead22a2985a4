"""Hyper-parameter defaults for the coefficient model -- the part of
pde_superresolution/training.py the integration path depends on
(create_hparams, training.py:40-165).  Training itself is out of scope."""
import copy
import json


class HParams(object):
  """Plain attribute bag standing in for tf.contrib.training.HParams."""

  def __init__(self, **kwargs):
    self.__dict__.update(kwargs)

  def override_from_dict(self, values):
    for name, value in values.items():
      if name not in self.__dict__:
        raise ValueError('Unknown hyperparameter: {}'.format(name))
      setattr(self, name, value)
    return self

  def set_hparam(self, name, value):
    """tf.contrib.training.HParams.set_hparam: the name must exist, the value is cast to the
    type of the current value (scripts/run_evaluation.py:131-132 overrides equation_kwargs this way)."""
    if name not in self.__dict__:
      raise KeyError('Hyperparameter {!r} does not exist'.format(name))
    current = self.__dict__[name]
    if isinstance(current, list):
      if not isinstance(value, (list, tuple)):
        raise ValueError('Must pass a list for multi-valued parameter: {}'.format(name))
      kind = type(current[0]) if current else float
      value = [self._cast(kind, v, name) for v in value]
    else:
      if isinstance(value, (list, tuple)):
        raise ValueError('Must not pass a list for single-valued parameter: {}'.format(name))
      value = self._cast(type(current), value, name)
    setattr(self, name, value)

  @staticmethod
  def _cast(kind, value, name):
    if kind is bool:
      if isinstance(value, str):
        if value.lower() in ('true', '1'):
          return True
        if value.lower() in ('false', '0'):
          return False
        raise ValueError('Could not parse {!r} as a bool for {}'.format(value, name))
      return bool(value)
    if kind is int and isinstance(value, float) and value != int(value):
      raise ValueError('Could not cast {!r} to int for {}'.format(value, name))
    return kind(value)

  def parse(self, text):
    """`name=value,name=[v1,v2],...` as accepted by the --hparams flag (scripts/run_training.py:47-51,73)."""
    import re
    pos, text = 0, text.strip()
    pattern = re.compile(r'\s*([A-Za-z_][A-Za-z0-9_]*)\s*=\s*(\[[^\]]*\]|[^,\[]*)\s*(?:,|$)')
    while pos < len(text):
      m = pattern.match(text, pos)
      if not m:
        raise ValueError('Malformed hyperparameter value: {}'.format(text[pos:]))
      pos = m.end()
      name, raw = m.group(1), m.group(2).strip()
      if raw.startswith('['):
        self.set_hparam(name, [v.strip() for v in raw[1:-1].split(',') if v.strip()])
      else:
        self.set_hparam(name, raw)
    return self

  def values(self):
    return copy.deepcopy(self.__dict__)

  def to_json(self):
    return json.dumps(self.values(), sort_keys=True)

  def __repr__(self):
    return 'HParams({})'.format(', '.join('%s=%r' % kv for kv in sorted(self.__dict__.items())))


_DEFAULTS = dict(
    # dataset (training.py:126-131)
    conservative=True, numerical_flux=False, equation_kwargs='{}', resample_factor=4,
    # network (training.py:132-141)
    model_target='coefficients', num_layers=3, filter_size=32, kernel_size=5, nonlinearity='relu',
    polynomial_accuracy_order=1, polynomial_accuracy_scale=1.0, ensure_unbiased_coefficients=False,
    coefficient_grid_min_size=6,
    # optimisation / noise / loss (training.py:142-163): carried for checkpoint
    # compatibility, unused by the integrator
    base_batch_size=128, learning_rates=[1e-3, 1e-4], learning_stops=[20000, 40000],
    frac_training=0.8, eval_interval=250, noise_probability=0.0, noise_amplitude=0.0,
    noise_type='white', ground_truth_order=-1, num_time_steps=0, error_floor_quantile=0.1,
    error_scale=[float('nan')], error_floor=[float('nan')], error_max=0.0,
    absolute_error_weight=1.0, relative_error_weight=0.0, space_derivatives_weight=0.0,
    time_derivative_weight=1.0, integrated_solution_weight=0.0,
)


def create_hparams(equation, **kwargs):
  """Defaults of training.create_hparams (training.py:125-165) with overrides."""
  hparams = HParams(equation=equation, **copy.deepcopy(_DEFAULTS))
  hparams.override_from_dict(kwargs)
  return hparams


def checkpoint_dir_to_path(checkpoint_dir):
  """training.py:523-524."""
  from . import checkpoint
  return checkpoint.checkpoint_dir_to_path(checkpoint_dir)


def load_hparams(checkpoint_dir):
  """Hyper-parameters saved next to a checkpoint as `hparams.pbtxt` (training.py:639-647);
  anything the file does not mention takes its create_hparams default."""
  import os
  from . import checkpoint
  with open(os.path.join(checkpoint_dir, 'hparams.pbtxt'), 'r') as f:
    values = checkpoint.parse_hparams_pbtxt(f.read())
  return create_hparams(**values)


def save_hparams(checkpoint_dir, hparams):
  """Write `hparams.pbtxt` the way training_loop does (training.py:590-592)."""
  import os
  from . import checkpoint
  os.makedirs(checkpoint_dir, exist_ok=True)
  with open(os.path.join(checkpoint_dir, 'hparams.pbtxt'), 'w') as f:
    f.write(checkpoint.format_hparams_pbtxt(hparams.values()))
