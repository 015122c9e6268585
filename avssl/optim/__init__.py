from .scheduler import get_scheduler
