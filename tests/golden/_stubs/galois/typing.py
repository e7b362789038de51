ElementLike = object
