from .accuracy import Accuracy
from .map import MAP
from .recall import Recall
