"""GGF move text / game record formatting used by the self-play worker (reference lib/ggf.py:56-100)."""
from datetime import datetime, timezone


def convert_action_to_move(action):
    """lib/ggf.py: 0 -> 'A1' ... 63 -> 'H8' (letter = row, digit = column+1); None -> 'PA'."""
    if action is None:
        return "PA"
    y, x = action // 8, action % 8
    return chr(ord("A") + y) + str(x + 1)


def convert_move_to_action(move_str):
    if move_str[:2].lower() == "pa":
        return None
    return (ord(move_str[0].upper()) - ord("A")) * 8 + int(move_str[1]) - 1


def make_ggf_string(black_name=None, white_name=None, dt=None, moves=None, result=None, think_time_sec=60):
    dt = dt or datetime.now(timezone.utc).replace(tzinfo=None)  # the reference's naive datetime.utcnow(): "%Z" prints nothing
    body = "".join(f"{'B' if i % 2 == 0 else 'W'}[{m}]" for i, m in enumerate(moves or []))
    return ("(;GM[Othello]PC[RAZSelf]DT[%s]PB[%s]PW[%s]RE[%s]TI[%d:%d]TY[8]"
            "BO[8 ---------------------------O*------*O--------------------------- *]%s;)") % (
        dt.strftime("%Y.%m.%d_%H:%M:%S.%Z"), black_name or "black", white_name or "white", result or "?",
        think_time_sec // 60, think_time_sec % 60, body)
