from .hyrax import limit
