# Intentionally empty
